// Training-side tail of the hot path (model.py:136-142): log_softmax over the classes and
// nn.CTCLoss(reduction='sum', zero_infinity=True) with blank 0, plus the gradient of that loss with respect to the raw
// decoder outputs ("probits"), i.e. what autograd hands to the decoder's backward.
//
//   k_row_lse          lse[r] = log sum_c exp(logits[r, c])            (one warp per row; HBM bound: reads M*C*4 bytes once)
//   k_log_softmax      out[r, c] = logits[r, c] - lse[r]               (A11; materialised only when the caller wants log-probs)
//   k_ctc_alpha        per line: alpha recursion in log space -> nll[b]; alpha kept for the backward pass
//   k_ctc_grad_init    grad[b, t, c] = softmax(logits)[b, t, c] for t < T_b and finite loss, else 0
//   k_ctc_beta_grad    per line: beta recursion; grad[b, t, l'_s] -= posterior of state s at t, summed per class in a FIXED
//                      order (chains of equal labels, blanks through a shuffle tree): no atomics, bit-reproducible
//
// Gradient identity: with lp = log_softmax(z), dL/dz[t, c] = softmax(z)[t, c] - (1/p(l|x)) sum_{s: l'_s = c} alpha_t(s) beta_t(s)
// / y_t(c) (Graves et al. 2006, eq. 16); this is exactly what torch returns for CTCLoss(log_softmax(z)) because the CTC
// backward's rows sum to zero and pass through log_softmax's backward unchanged.
#include <math.h>

#include "common.h"

namespace cocr {

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float log_add(float a, float b) {           // log(exp(a) + exp(b)), -inf safe
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return -INFINITY;
    return m + log1pf(expf(fminf(a, b) - m));
}

// one warp per row; float4 loads when the row is 16-byte aligned
__global__ void __launch_bounds__(256)
k_row_lse(const float *__restrict__ logits, int64_t rows, int C, float *__restrict__ lse) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *p = logits + row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, p[c]);
    m = warp_max(m);
    float s = 0.f;
    if (m > -INFINITY)
        for (int c = lane; c < C; c += 32) s += expf(p[c] - m);
    s = warp_sum(s);
    if (lane == 0) lse[row] = (m > -INFINITY) ? m + logf(s) : -INFINITY;
}

__global__ void __launch_bounds__(256)
k_log_softmax(const float *__restrict__ logits, const float *__restrict__ lse, int64_t rows, int C, float *__restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float l = lse[row];
    const float *p = logits + row * C;
    float *o = out + row * C;
    for (int c = lane; c < C; c += 32) o[c] = p[c] - l;
}

// Extended label sequence of line b: l'_s = blank (0) for even s, target[(s-1)/2] for odd s; S = 2 L + 1.
struct CtcLine {
    int Tn, Ln, S;
    const int32_t *tgt;
};
__device__ __forceinline__ CtcLine ctc_line(const int32_t *targets, const int32_t *target_lens, const int32_t *input_lens, int b, int T,
                                            int max_target_len) {
    CtcLine l;
    int off = 0;
    for (int i = 0; i < b; ++i) off += max(target_lens[i], 0);
    l.Ln = min(max(target_lens[b], 0), max_target_len);
    l.Tn = input_lens ? min(max(input_lens[b], 0), T) : T;
    l.S = 2 * l.Ln + 1;
    l.tgt = targets + off;
    return l;
}

// Frames of a line staged per chunk in shared memory: the recursions are chains of T dependent steps, and a step that waits for
// global memory (its emission gather, the row the previous step wrote) costs a microsecond; from shared memory it costs ~0.1.
__host__ __device__ inline int ctc_chunk(int Smax, int rows_per_frame) {
    int tc = (64 * 1024) / (Smax * 4 * rows_per_frame);
    return tc < 1 ? 1 : (tc > 32 ? 32 : tc);
}

// alpha recursion.  alpha: (B, T, Smax) global scratch (kept for the backward pass); nll: (B) = -log p(l | x), +inf when the
// target cannot be emitted in T_b frames.  One block per line, thread = state (strided).
// Dynamic shared memory: em[TC][Smax] emissions of the chunk | a[2][Smax] | lab[Smax] | skip[Smax].
__global__ void __launch_bounds__(256)
k_ctc_alpha(const float *__restrict__ logits, const float *__restrict__ lse, const int32_t *__restrict__ targets,
            const int32_t *__restrict__ target_lens, const int32_t *__restrict__ input_lens, int T, int C, int max_target_len, int Smax,
            float *__restrict__ alpha, float *__restrict__ nll) {
    extern __shared__ float sm[];
    const int TC = ctc_chunk(Smax, 1);
    float *em = sm, *a0 = sm + (size_t)TC * Smax, *a1 = a0 + Smax;
    int *lab = reinterpret_cast<int *>(a1 + Smax), *skip = lab + Smax;
    const int b = blockIdx.x;
    const CtcLine ln = ctc_line(targets, target_lens, input_lens, b, T, max_target_len);
    const float *lg = logits + (int64_t)b * T * C;
    const float *ls = lse + (int64_t)b * T;
    float *al = alpha + (int64_t)b * T * Smax;
    if (ln.Tn == 0) {
        if (threadIdx.x == 0) nll[b] = ln.Ln == 0 ? 0.0f : INFINITY;
        return;
    }
    for (int s = threadIdx.x; s < ln.S; s += blockDim.x) {
        const int l = (s & 1) ? ln.tgt[s >> 1] : 0;
        lab[s] = l;
        skip[s] = (s >= 2 && (s & 1) && l != ln.tgt[(s >> 1) - 1]) ? 1 : 0;
    }
    __syncthreads();
    for (int t0 = 0; t0 < ln.Tn; t0 += TC) {
        const int nt = min(TC, ln.Tn - t0);
        for (int e = threadIdx.x; e < nt * ln.S; e += blockDim.x) {         // the chunk's emissions, gathered in parallel
            const int tt = e / ln.S, s_ = e - tt * ln.S;
            em[tt * Smax + s_] = lg[(int64_t)(t0 + tt) * C + lab[s_]] - ls[t0 + tt];
        }
        __syncthreads();
        for (int tt = 0; tt < nt; ++tt) {
            const int t = t0 + tt;
            const float *prev = (t & 1) ? a0 : a1;
            float *cur = (t & 1) ? a1 : a0;
            for (int s = threadIdx.x; s < ln.S; s += blockDim.x) {
                float a;
                if (t == 0) {
                    a = (s <= 1) ? em[s] : -INFINITY;
                } else {
                    a = prev[s];
                    if (s >= 1) a = log_add(a, prev[s - 1]);
                    if (skip[s]) a = log_add(a, prev[s - 2]);
                    a += em[tt * Smax + s];
                }
                cur[s] = a;
                al[(int64_t)t * Smax + s] = a;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const float *last = ((ln.Tn - 1) & 1) ? a1 : a0;
        const float tail = ln.S == 1 ? last[0] : log_add(last[ln.S - 1], last[ln.S - 2]);
        nll[b] = -tail;
    }
}

// loss_sum = sum_b (finite nll_b ? nll_b : 0)   (reduction='sum', zero_infinity=True); fixed summation order
__global__ void k_ctc_reduce(const float *__restrict__ nll, int B, float *__restrict__ loss_sum, float *__restrict__ per_line) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        const float v = isfinite(nll[b]) ? nll[b] : 0.0f;
        if (per_line) per_line[b] = v;
        s += v;
    }
    *loss_sum = s;
}

// dense part of the gradient: softmax rows for the valid frames of lines with a finite loss, zeros elsewhere
__global__ void __launch_bounds__(256)
k_ctc_grad_init(const float *__restrict__ logits, const float *__restrict__ lse, const float *__restrict__ nll,
                const int32_t *__restrict__ input_lens, int64_t rows, int T, int C, float *__restrict__ grad) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = (int)(row / T), t = (int)(row - (int64_t)b * T);
    const int Tn = input_lens ? min(max(input_lens[b], 0), T) : T;
    const bool live = t < Tn && isfinite(nll[b]);
    const float l = lse[row];
    const float *p = logits + row * C;
    float *g = grad + row * C;
    for (int c = lane; c < C; c += 32) g[c] = live ? expf(p[c] - l) : 0.0f;
}

// beta recursion + posterior subtraction.  Dynamic shared memory: em[TC][Smax] | alc[TC][Smax] (emissions and alpha rows of the chunk
// of frames being walked, loaded in parallel; alc is overwritten with the per-class posterior sums of the frame, which are
// subtracted from the gradient in parallel after the chunk) | beta[2][Smax] | post[Smax] | nxt[Smax] | first[Smax] | lab[Smax] | skip[Smax] (int).
//   nxt[s] (odd s): next odd state with the same label, or -1; `first` states own their class and sum their chain in order.
__global__ void __launch_bounds__(256)
k_ctc_beta_grad(const float *__restrict__ logits, const float *__restrict__ lse, const int32_t *__restrict__ targets,
                const int32_t *__restrict__ target_lens, const int32_t *__restrict__ input_lens, int T, int C, int max_target_len, int Smax,
                const float *__restrict__ alpha, const float *__restrict__ nll, float *__restrict__ grad) {
    extern __shared__ float sm[];
    const int TC = ctc_chunk(Smax, 2);
    float *em = sm, *alc = em + (size_t)TC * Smax, *beta0 = alc + (size_t)TC * Smax, *beta1 = beta0 + Smax, *post = beta1 + Smax;
    int *nxt = reinterpret_cast<int *>(post + Smax);
    int *first = nxt + Smax, *lab = first + Smax, *skip = lab + Smax;
    const int b = blockIdx.x;
    const CtcLine ln = ctc_line(targets, target_lens, input_lens, b, T, max_target_len);
    const float nl = nll[b];
    if (ln.Tn == 0 || !isfinite(nl)) return;                       // zero_infinity: the gradient of this line stays zero
    const float *lg = logits + (int64_t)b * T * C;
    const float *ls = lse + (int64_t)b * T;
    const float *al = alpha + (int64_t)b * T * Smax;
    float *gr = grad + (int64_t)b * T * C;
    // chains of equal labels (fixed order -> deterministic sums)
    for (int s = threadIdx.x; s < ln.S; s += blockDim.x) {
        nxt[s] = -1;
        first[s] = 0;
        const int l = (s & 1) ? ln.tgt[s >> 1] : 0;
        lab[s] = l;
        skip[s] = (s + 2 < ln.S && (s & 1) && l != ln.tgt[(s >> 1) + 1]) ? 1 : 0;
        if (s & 1) {
            int f = 1;
            for (int u = 1; u < s; u += 2)
                if (ln.tgt[u >> 1] == l) { f = 0; break; }
            first[s] = f;
            for (int u = s + 2; u < ln.S; u += 2)
                if (ln.tgt[u >> 1] == l) { nxt[s] = u; break; }
        }
    }
    __syncthreads();
    for (int thi = ln.Tn; thi > 0; thi -= TC) {                    // chunks of frames [tlo, thi), walked backwards
        const int tlo = max(0, thi - TC), nt = thi - tlo;
        for (int e = threadIdx.x; e < nt * ln.S; e += blockDim.x) {
            const int tt = e / ln.S, s_ = e - tt * ln.S;
            em[tt * Smax + s_] = lg[(int64_t)(tlo + tt) * C + lab[s_]] - ls[tlo + tt];
            alc[tt * Smax + s_] = al[(int64_t)(tlo + tt) * Smax + s_];
        }
        __syncthreads();
        for (int tt = nt - 1; tt >= 0; --tt) {
            const int t = tlo + tt;
            float *cur = ((ln.Tn - 1 - t) & 1) ? beta1 : beta0;
            const float *nx = ((ln.Tn - 1 - t) & 1) ? beta0 : beta1;
            for (int s = threadIdx.x; s < ln.S; s += blockDim.x) {
                const float lp = em[tt * Smax + s];
                float bt;
                if (t == ln.Tn - 1) {
                    bt = (s >= ln.S - 2) ? lp : -INFINITY;         // the last blank and the last label may end the path
                } else {
                    bt = nx[s];
                    if (s + 1 < ln.S) bt = log_add(bt, nx[s + 1]);
                    if (skip[s]) bt = log_add(bt, nx[s + 2]);
                    bt += lp;
                }
                cur[s] = bt;
                // posterior of state s at frame t: alpha_t(s) beta_t(s) / (y_t(l'_s) p(l | x))   (beta includes y_t here)
                const float lpost = alc[tt * Smax + s] + bt - lp + nl;
                post[s] = (lpost > -INFINITY) ? expf(lpost) : 0.0f;
            }
            __syncthreads();
            // (the alpha row of this frame has been consumed above: its slots now take the frame's per-class sums, applied after the chunk
            // so that the read-modify-write of the gradient is not part of the sequential chain)
            // non-blank classes: the first state of every label sums its chain
            for (int s = 1 + 2 * (int)threadIdx.x; s < ln.S; s += 2 * blockDim.x) {
                if (!first[s]) continue;
                float acc = 0.f;
                for (int u = s; u >= 0; u = nxt[u]) acc += post[u];
                alc[tt * Smax + s] = acc;
            }
            // blank: warp 0, strided partial sums in a fixed order + shuffle tree
            if (threadIdx.x < 32) {
                float acc = 0.f;
                for (int s = 2 * (int)threadIdx.x; s < ln.S; s += 64) acc += post[s];
                acc = warp_sum(acc);
                if (threadIdx.x == 0) alc[tt * Smax] = acc;
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < nt * ln.S; e += blockDim.x) {          // one (frame, class) per element: no two touch the same address
            const int tt = e / ln.S, s_ = e - tt * ln.S;
            if (s_ == 0 || first[s_]) gr[(int64_t)(tlo + tt) * C + lab[s_]] -= alc[tt * Smax + s_];
        }
        __syncthreads();
    }
}

// ---- decoder backward (nn.Linear(d, C), pred.py:90 / model.py:114): dW = G^T E, db = colsum(G), dE = G W ----------------------
// out[z][c][j] = sum over the z-th slice of rows m of G[m][c] * E1[m][j], with E1 = [E | 1] (column d carries the bias gradient).
// 32 x 32 output tiles, 32 rows of m per shared-memory step; partial sums per slice are reduced in slice order by
// k_dec_reduce (fixed order: bit-reproducible, no atomics).
constexpr int DB_T = 32;
__global__ void __launch_bounds__(256)
k_dec_wgrad_partial(const float *__restrict__ G, const float *__restrict__ E, int M, int C, int d, int rows_per_slice,
                    float *__restrict__ partial /* [Z][C][d + 1] */) {
    __shared__ float sg[DB_T][DB_T + 1], se[DB_T][DB_T + 1];
    const int c0 = blockIdx.x * DB_T, j0 = blockIdx.y * DB_T, z = blockIdx.z;
    const int m_lo = z * rows_per_slice, m_hi = min(M, m_lo + rows_per_slice);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;               // 32 x 8 threads; thread -> (c = c0 + ty*4 + i, j = j0 + tx)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m0 = m_lo; m0 < m_hi; m0 += DB_T) {
        for (int r = ty; r < DB_T; r += 8) {
            const int m = m0 + r;
            sg[r][tx] = (m < m_hi && c0 + tx < C) ? G[(int64_t)m * C + c0 + tx] : 0.0f;
            const int j = j0 + tx;
            se[r][tx] = (m < m_hi) ? (j < d ? E[(int64_t)m * d + j] : (j == d ? 1.0f : 0.0f)) : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < DB_T; ++r) {
            const float e = se[r][tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(sg[r][ty * 4 + i], e, acc[i]);
        }
        __syncthreads();
    }
    const int j = j0 + tx;
    if (j <= d)
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty * 4 + i;
            if (c < C) partial[((int64_t)z * C + c) * (d + 1) + j] = acc[i];
        }
}
__global__ void k_dec_reduce(const float *__restrict__ partial, int Z, int C, int d, float *__restrict__ gw, float *__restrict__ gb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * (d + 1)) return;
    float s = 0.f;
    for (int z = 0; z < Z; ++z) s += partial[(int64_t)z * C * (d + 1) + i];
    const int c = i / (d + 1), j = i - c * (d + 1);
    if (j < d) gw[(int64_t)c * d + j] = s;
    else gb[c] = s;
}
// dE[m][j] = sum_c G[m][c] W[c][j]; one warp per row, W streamed through L1/L2 (C * d * 4 bytes, small)
__global__ void __launch_bounds__(256)
k_dec_dinput(const float *__restrict__ G, const float *__restrict__ W, int M, int C, int d, float *__restrict__ dE) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *g = G + row * C;
    for (int j0 = 0; j0 < d; j0 += 128) {                                // 4 output columns per lane per sweep
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) {
            const float gv = g[c];                                       // warp-uniform broadcast
            const float *w = W + (int64_t)c * d + j0 + lane;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j0 + lane + 32 * i < d) a[i] = fmaf(gv, w[32 * i], a[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (j0 + lane + 32 * i < d) dE[row * d + j0 + lane + 32 * i] = a[i];
    }
}

}  // namespace

int launch_decoder_backward(cocr_handle *h, const float *G, const float *E, int M, cudaStream_t st, float *gw, float *gb, float *dE) {
    const int C = h->dm.C, d = h->dm.d;
    int Z = (M + 1023) / 1024;
    if (Z > 64) Z = 64;
    const int rows_per_slice = ((M + Z - 1) / Z + DB_T - 1) / DB_T * DB_T;
    float *partial = nullptr;
    COCR_CUDA(h, cudaMallocAsync((void **)&partial, (size_t)Z * C * (d + 1) * 4, st));
    int rc = COCR_OK;
    do {
        {
            ProfScope ps(h, "decoder_wgrad", st, 2.0 * M * (double)C * (d + 1), 4.0 * M * (double)(C + d));
            dim3 grid((C + DB_T - 1) / DB_T, (d + 1 + DB_T - 1) / DB_T, Z);
            k_dec_wgrad_partial<<<grid, 256, 0, st>>>(G, E, M, C, d, rows_per_slice, partial);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_dec_wgrad_partial")) != COCR_OK) break;
            h->launches++;
            k_dec_reduce<<<(C * (d + 1) + 255) / 256, 256, 0, st>>>(partial, Z, C, d, gw, gb);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_dec_reduce")) != COCR_OK) break;
            h->launches++;
        }
        if (dE) {
            ProfScope ps(h, "decoder_dinput", st, 2.0 * M * (double)C * d, 4.0 * M * (double)(C + d));
            k_dec_dinput<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(G, h->dec_w, M, C, d, dE);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_dec_dinput")) != COCR_OK) break;
            h->launches++;
        }
    } while (0);
    cudaFreeAsync(partial, st);
    return rc;
}

int launch_log_softmax(cocr_handle *h, const float *logits, int64_t rows, int C, cudaStream_t st, float *out) {
    if (rows <= 0) return COCR_OK;
    float *lse = nullptr;
    COCR_CUDA(h, cudaMallocAsync((void **)&lse, (size_t)rows * 4, st));
    const unsigned grid = (unsigned)((rows + 7) / 8);
    {
        ProfScope ps(h, "row_logsumexp", st, 0.0, (double)rows * C * 4.0);
        k_row_lse<<<grid, 256, 0, st>>>(logits, rows, C, lse);
        COCR_LAUNCHED(h);
    }
    {
        ProfScope ps(h, "log_softmax", st, 0.0, (double)rows * C * 8.0);
        k_log_softmax<<<grid, 256, 0, st>>>(logits, lse, rows, C, out);
        COCR_LAUNCHED(h);
    }
    cudaFreeAsync(lse, st);
    return COCR_OK;
}

int launch_ctc_loss(cocr_handle *h, const float *logits, const int32_t *targets, const int32_t *target_lens, const int32_t *input_lens,
                    int B, int T, int C, int max_target_len, cudaStream_t st, float *loss_sum, float *loss_per_line, float *grad) {
    const int64_t rows = (int64_t)B * T;
    const int Smax = 2 * max_target_len + 1;
    float *lse = nullptr, *alpha = nullptr, *nll = nullptr;
    COCR_CUDA(h, cudaMallocAsync((void **)&lse, (size_t)rows * 4, st));
    COCR_CUDA(h, cudaMallocAsync((void **)&alpha, (size_t)rows * Smax * 4, st));
    COCR_CUDA(h, cudaMallocAsync((void **)&nll, (size_t)B * 4, st));
    const unsigned grid = (unsigned)((rows + 7) / 8);
    int rc = COCR_OK;
    do {
        {
            ProfScope ps(h, "row_logsumexp", st, 0.0, (double)rows * C * 4.0);
            k_row_lse<<<grid, 256, 0, st>>>(logits, rows, C, lse);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_row_lse")) != COCR_OK) break;
            h->launches++;
        }
        {
            ProfScope ps(h, "ctc_alpha", st, 0.0, (double)rows * Smax * 4.0);
            const size_t asmem = ((size_t)ctc_chunk(Smax, 1) * Smax + (size_t)4 * Smax) * 4;
            if (asmem > 48 * 1024) {
                static PerDevice<size_t> aattr;
                if (asmem > aattr.get()) {
                    if ((rc = check_cuda(h, cudaFuncSetAttribute(k_ctc_alpha, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asmem),
                                         "k_ctc_alpha shared memory")) != COCR_OK) break;
                    aattr.get() = asmem;
                }
            }
            k_ctc_alpha<<<B, 256, asmem, st>>>(logits, lse, targets, target_lens, input_lens, T, C, max_target_len, Smax, alpha, nll);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_ctc_alpha")) != COCR_OK) break;
            h->launches++;
        }
        k_ctc_reduce<<<1, 32, 0, st>>>(nll, B, loss_sum, loss_per_line);
        if ((rc = check_cuda(h, cudaGetLastError(), "k_ctc_reduce")) != COCR_OK) break;
        h->launches++;
        if (grad) {
            {
                ProfScope ps(h, "ctc_grad_softmax", st, 0.0, (double)rows * C * 8.0);
                k_ctc_grad_init<<<grid, 256, 0, st>>>(logits, lse, nll, input_lens, rows, T, C, grad);
                if ((rc = check_cuda(h, cudaGetLastError(), "k_ctc_grad_init")) != COCR_OK) break;
                h->launches++;
            }
            const size_t smem = ((size_t)2 * ctc_chunk(Smax, 2) * Smax + (size_t)7 * Smax) * 4;
            if (smem > 48 * 1024) {
                static PerDevice<size_t> attr;
                if (smem > attr.get()) {
                    if ((rc = check_cuda(h, cudaFuncSetAttribute(k_ctc_beta_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                         "k_ctc_beta_grad shared memory")) != COCR_OK) break;
                    attr.get() = smem;
                }
            }
            ProfScope ps(h, "ctc_beta_grad", st, 0.0, (double)rows * Smax * 4.0);
            k_ctc_beta_grad<<<B, 256, smem, st>>>(logits, lse, targets, target_lens, input_lens, T, C, max_target_len, Smax, alpha, nll, grad);
            if ((rc = check_cuda(h, cudaGetLastError(), "k_ctc_beta_grad")) != COCR_OK) break;
            h->launches++;
        }
    } while (0);
    cudaFreeAsync(lse, st);
    cudaFreeAsync(alpha, st);
    cudaFreeAsync(nll, st);
    return rc;
}

}  // namespace cocr
