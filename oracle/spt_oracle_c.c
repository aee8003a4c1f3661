/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C restatement of the integer / order-sensitive parts of the SPT hot path
 * (reference: ytgui/SPT-proto, cited as file:line relative to /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (spt_proto_b200/) never does.
 *
 * Functions
 *   oracle_lookup_literal : thread-by-thread emulation of lookup_forward_kernel
 *                           (extension/lookup.cu:11-84), generalised to any
 *                           (n_subspaces >= 4, nonzeros % 4 == 0).
 *   oracle_cdist_forward  : sequential-fp32 L1 distance + strict-< running argmin
 *                           (extension/cdist.cu:28-68).
 *   oracle_csr2csc        : stable counting-sort transpose of a shared-indptr CSR
 *                           (semantics of legacy/csr2csc.cpp:3-54 / cusparseCsr2cscEx2).
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC, no fast-math so the fp32
 * summation order below is preserved).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_SIZE 16  /* lookup.cu:4  */
#define WORKER_SIZE 4  /* lookup.cu:5  */
#define N_SLOTS 4      /* lookup.cu:31 */

static inline int imin(int a, int b) { return a < b ? a : b; }

/*
 * One row (b, gy) of lookup_forward_kernel.  Rows of a block only interact through
 * __syncthreads(), every shared array is indexed by the row's own ty, so rows can be
 * emulated independently.  Lanes tx=0..3 of a row sit in the same warp and run the
 * `for local_x = tx; ...; local_x += WORKER_SIZE` loop (lookup.cu:48) in lock-step:
 * iteration `step` of all four lanes is one instruction.  The only hardware-undefined case is
 * two lanes storing to the same shared address in the same instruction (lane 3 / 2 overflowing
 * onto lane 0 / 1's last slot while that lane writes it).  MEASURED on B200 (sm_100a) against
 * the compiled reference kernel (oracle/_ref/ext_ref.so, tests/test_kernels_gpu.py::
 * test_lookup_vs_reference_kernel): the LOWEST lane's store survives, run after run.  We
 * therefore serialise each instruction as tx = 3,2,1,0 (last writer = lowest lane).  SURVEY.md
 * section 8 a-2 guessed "higher lane wins" without a GPU; the measurement overrides it.
 */
static void lookup_row(const int32_t *lhs_row, const int32_t *rhs, int32_t *out_row,
                       int seq_length, int n_spaces, int nnz, int gy, uint16_t *ind /* [N_SLOTS][nnz] */) {
    int cursors[WORKER_SIZE][N_SLOTS];
    for (int tx = 0; tx < WORKER_SIZE; ++tx)
        for (int s = 0; s < N_SLOTS; ++s) cursors[tx][s] = tx; /* lookup.cu:31 */
    const int block_row0 = gy - (gy % BLOCK_SIZE);            /* gy - ty */
    const int div = n_spaces / N_SLOTS;                        /* lookup.cu:62 */
    for (int offset_x = 0; offset_x < seq_length; offset_x += BLOCK_SIZE) {
        if (offset_x > block_row0) break;                      /* lookup.cu:36-38 */
        for (int step = 0; step < BLOCK_SIZE / WORKER_SIZE; ++step) {
            for (int tx = WORKER_SIZE - 1; tx >= 0; --tx) {
                int local_x = tx + step * WORKER_SIZE;
                int j = offset_x + local_x;
                if (j > gy) continue;                          /* tril, lookup.cu:50-52 (break == continue: j only grows) */
                int count = 0;
                for (int k = 0; k < n_spaces; ++k)             /* uint16 caches, lookup.cu:22,40 */
                    count += ((uint16_t)lhs_row[k] == (uint16_t)rhs[(size_t)j * n_spaces + k]);
                int slot = imin(N_SLOTS - 1, count / div);     /* lookup.cu:61-63 */
                int cursor = cursors[tx][slot];
                ind[slot * nnz + cursor] = (uint16_t)j;        /* lookup.cu:65 */
                cursors[tx][slot] = imin(cursor + WORKER_SIZE, nnz - tx - 1); /* lookup.cu:66 */
            }
        }
    }
    /* store, lookup.cu:71-83 */
    for (int tx = 0; tx < WORKER_SIZE; ++tx) {
        int slot = N_SLOTS - 1, cursor = tx;
        int lim = imin(gy + 1, nnz);
        for (int local_x = tx; local_x < lim; local_x += WORKER_SIZE) {
            while (slot >= 0 && cursor >= cursors[tx][slot]) {
                slot -= 1;
                cursor = tx;
            }
            if (slot < 0) break;
            out_row[local_x] = (int32_t)ind[slot * nnz + cursor];
            cursor += WORKER_SIZE;
        }
    }
}

/* left/right: [batch, seq, n_spaces] int32; output: [batch, seq, nnz] int32 (zero-filled here,
 * torch::zeros at lookup.cu:107).  Returns 0 on success, 1 on unsupported shape. */
int oracle_lookup_literal(const int32_t *left, const int32_t *right, int32_t *output,
                          int batch, int seq_length, int n_spaces, int nnz) {
    if (n_spaces < N_SLOTS || nnz < WORKER_SIZE || seq_length % BLOCK_SIZE != 0) return 1;
    memset(output, 0, (size_t)batch * seq_length * nnz * sizeof(int32_t));
    uint16_t *ind = (uint16_t *)malloc((size_t)N_SLOTS * nnz * sizeof(uint16_t));
    if (!ind) return 2;
    for (int b = 0; b < batch; ++b) {
        const int32_t *rhs = right + (size_t)b * seq_length * n_spaces;
        for (int gy = 0; gy < seq_length; ++gy) {
            memset(ind, 0, (size_t)N_SLOTS * nnz * sizeof(uint16_t));
            lookup_row(left + ((size_t)b * seq_length + gy) * n_spaces, rhs,
                       output + ((size_t)b * seq_length + gy) * nnz, seq_length, n_spaces, nnz, gy, ind);
        }
    }
    free(ind);
    return 0;
}

/* query [m, n, dc], table [m, c, dc] -> distance [m, n, c] (may be NULL), indices [m, n].
 * Sum over i ascending in fp32, strict '<' against a running minimum that starts at 1e13
 * (cdist.cu:29,46-54) => lowest codeword index wins ties. */
int oracle_cdist_forward(const float *query, const float *table, float *distance, int32_t *indices,
                         int m, int n, int c, int dc) {
    for (int s = 0; s < m; ++s) {
        for (int q = 0; q < n; ++q) {
            const float *qv = query + ((size_t)s * n + q) * dc;
            int min_index = 0;
            float min_distance = 1e13f;
            for (int w = 0; w < c; ++w) {
                const float *tv = table + ((size_t)s * c + w) * dc;
                volatile float reduced = 0.0f; /* volatile: forbid re-association / vector reduction */
                for (int i = 0; i < dc; ++i) reduced = reduced + fabsf(qv[i] - tv[i]);
                float r = reduced;
                if (r < min_distance) {
                    min_distance = r;
                    min_index = w;
                }
                if (distance) distance[((size_t)s * n + q) * c + w] = r;
            }
            indices[(size_t)s * n + q] = min_index;
        }
    }
    return 0;
}

/* CSR (shared indptr [S+1], per-batch indices/values [B, nnz]) -> CSC with a shared? No: column
 * populations differ per batch item, so the CSC has per-batch col pointers.
 *   col_ptr  [B, n_cols+1]
 *   row_idx  [B, nnz]   row of every entry, grouped by column, rows ascending inside a column
 *                        (stable: entries of the same (row, col) keep their CSR order)
 *   perm     [B, nnz]   position of the entry in the CSR value array (so values_csc = values[perm]) */
int oracle_csr2csc(const int32_t *indptr, const int32_t *indices, int32_t *col_ptr, int32_t *row_idx,
                   int32_t *perm, int batch, int n_rows, int n_cols, int nnz) {
    int32_t *cursor = (int32_t *)malloc((size_t)(n_cols + 1) * sizeof(int32_t));
    if (!cursor) return 2;
    for (int b = 0; b < batch; ++b) {
        const int32_t *idx = indices + (size_t)b * nnz;
        int32_t *cp = col_ptr + (size_t)b * (n_cols + 1);
        memset(cp, 0, (size_t)(n_cols + 1) * sizeof(int32_t));
        for (int e = 0; e < nnz; ++e) {
            if (idx[e] < 0 || idx[e] >= n_cols) { free(cursor); return 1; }
            cp[idx[e] + 1] += 1;
        }
        for (int c = 0; c < n_cols; ++c) cp[c + 1] += cp[c];
        memcpy(cursor, cp, (size_t)(n_cols + 1) * sizeof(int32_t));
        for (int r = 0; r < n_rows; ++r) {
            for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
                int dst = cursor[idx[e]]++;
                row_idx[(size_t)b * nnz + dst] = r;
                perm[(size_t)b * nnz + dst] = e;
            }
        }
    }
    free(cursor);
    return 0;
}
