"""Temporarily instrument attn_bwd_kv2_tc_kernel with clock64 counters (debug only; restore the file afterwards)."""
p='/root/repo/spt_proto_b200/csrc/attn_tc.cu'
s=open(p).read()
i=s.index("attn_bwd_kv2_tc_kernel(const __grid_constant__")
head,tail=s[:i],s[i:]
def rep(a,b):
    global tail
    assert a in tail, a[:60]
    tail=tail.replace(a,b,1)
rep('''        mbar_wait(own_full, 0);
        issue_scores(0);
        for (int j = 0; j < n_tiles; ++j) {
            if (j + 1 < n_tiles) issue_scores(j + 1);''','''        long long d_t0 = clock64(), d_is = 0, d_wp = 0, d_ac = 0;
        mbar_wait(own_full, 0);
        long long d_own = clock64() - d_t0;
        issue_scores(0);
        for (int j = 0; j < n_tiles; ++j) {
            long long a_ = clock64();
            if (j + 1 < n_tiles) issue_scores(j + 1);
            d_is += clock64() - a_;''')
rep('''            mbar_wait(p_full(j & 1), (j >> 1) & 1);
            fence_after_sync();
            if (elect_one()) {
                const uint64_t off = (uint64_t)((st * T_BYTES) >> 4);''','''            long long b_ = clock64();
            mbar_wait(p_full(j & 1), (j >> 1) & 1);
            d_wp += clock64() - b_;
            fence_after_sync();
            long long c_ = clock64();
            if (elect_one()) {
                const uint64_t off = (uint64_t)((st * T_BYTES) >> 4);''')
rep('''                if (j + 1 == n_tiles) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {''','''                if (j + 1 == n_tiles) umma_commit(acc_full);
            }
            __syncwarp();
            d_ac += clock64() - c_;
        }
        if (blockIdx.x == 4 && blockIdx.y == 7 && lane == 0) printf("kv issuer: total %lld own_wait %lld issue_scores %lld wait_p %lld issue_acc %lld tiles %d\\n", clock64() - d_t0, d_own, d_is, d_wp, d_ac, n_tiles);
    } else {''')
rep('''            mbar_wait(qd_full(st), (j / ST) & 1);
            mbar_wait(sc_full(j & 1), (j >> 1) & 1);
            fence_after_sync();''','''            long long a_ = clock64();
            mbar_wait(qd_full(st), (j / ST) & 1);
            mbar_wait(sc_full(j & 1), (j >> 1) & 1);
            fence_after_sync();
            long long b_ = clock64(); d_ws += b_ - a_;''')
rep('''        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % ST;
            const unsigned char *slot = rowq + st * KV2_ROWQ_BYTES;''','''        long long d_t0 = clock64(), d_ws = 0, d_ld = 0, d_m = 0, d_st = 0;
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % ST;
            const unsigned char *slot = rowq + st * KV2_ROWQ_BYTES;''')
rep('''            tmem_ld_wait();
            uint32_t pe[8], pd[8];''','''            tmem_ld_wait();
            long long c_ = clock64(); d_ld += c_ - b_;
            uint32_t pe[8], pd[8];''')
rep('''            tmem_st8(col, pe);
            tmem_st8(col + 32, pd);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(j & 1));
        }''','''            long long e_ = clock64(); d_m += e_ - c_;
            tmem_st8(col, pe);
            tmem_st8(col + 32, pd);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(p_full(j & 1));
            d_st += clock64() - e_;
        }
        if (blockIdx.x == 4 && blockIdx.y == 7 && lane == 0 && (warp == 0 || warp == 13)) printf("kv math warp %d: total %lld wait_sc %lld tmem_ld %lld math %lld st+arrive %lld\\n", warp, clock64() - d_t0, d_ws, d_ld, d_m, d_st);''')
open(p,'w').write(head+tail)
