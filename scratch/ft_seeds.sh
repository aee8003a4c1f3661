#!/bin/bash
# fine-tuning step on ONE GPU with the batches of ranks 0, 1, 2, 7: how much of the multi-GPU excess is data
for sd in 1234 1235 1236 1241; do
  python scripts/finetune_step.py --graph --data-seed $sd 2>/dev/null | tail -1 > gpurun_out/ft_seed_$sd.json
  python -c "import json; d=json.load(open('gpurun_out/ft_seed_$sd.json')); print('seed', $sd, round(d['ms_per_step'], 3))"
done
