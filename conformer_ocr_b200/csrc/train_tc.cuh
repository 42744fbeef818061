// Tensor-core variant of the training-step modules (included by train.cu inside its anonymous namespace): every dense
// contraction runs on tcgen05 through the batched GEMM of bgemm_tc.cu with bf16 operands and fp32 accumulation; parameters,
// the residual stream, LayerNorm / BatchNorm statistics, softmax, the loss and all gradients sums stay fp32 ("fp32 master,
// bf16 compute", SURVEY.md 8d cfg 4).  The arithmetic is the reference's own formulation (attention.py:82-101): content
// scores (q+u)k^T and positional scores (q+v)p^T over the 2T-1 relative offsets, the relative shift read as a contiguous row
// segment, so the backward pass has the same structure as the fp32 path and as autograd's.
#pragma once

using bf16 = __nv_bfloat16;

struct TcBuf {
    bf16 *pb, *peb, *xnb, *zb, *actb, *dyb, *wideb, *ctxb, *encb, *glogb, *dposb, *qu, *qv, *kb, *vb, *dctxb, *posb, *dSb, *dPSb;
    float *Sc, *PS, *delta;
};

inline TcBuf tc_buffers(char *ws, const Plan &p) {
    TcBuf t;
    auto H = [&](size_t off) { return (bf16 *)(ws + off); };
    auto F = [&](size_t off) { return (float *)(ws + off); };
    t.pb = H(p.pb); t.peb = H(p.peb); t.xnb = H(p.xnb); t.zb = H(p.zb); t.actb = H(p.actb); t.dyb = H(p.dyb); t.wideb = H(p.wideb);
    t.ctxb = H(p.ctxb); t.encb = H(p.encb); t.glogb = H(p.glogb); t.dposb = H(p.dposb); t.qu = H(p.qu); t.qv = H(p.qv); t.kb = H(p.kb);
    t.vb = H(p.vb); t.dctxb = H(p.dctxb); t.posb = H(p.posb); t.dSb = H(p.dSb); t.dPSb = H(p.dPSb);
    t.Sc = F(p.Sc); t.PS = F(p.PS); t.delta = F(p.delta);
    return t;
}

// operands saved per layer by the forward pass
struct TcSaved {
    bf16 *xnb[4], *actb[2], *zb, *ctxb, *qu, *qv, *kb, *vb, *posb;
};
inline TcSaved tc_saved(const TcBuf &t, const Plan &p, const Dims &dm, int l) {
    TcSaved v;
    const size_t M = p.M, d = dm.d, ff = dm.ff;
    for (int i = 0; i < 4; ++i) v.xnb[i] = t.xnb + ((size_t)l * 4 + i) * M * d;
    for (int i = 0; i < 2; ++i) v.actb[i] = t.actb + ((size_t)l * 2 + i) * M * ff;
    v.zb = t.zb + (size_t)l * M * d;
    v.ctxb = t.ctxb + (size_t)l * M * d;
    v.qu = t.qu + (size_t)l * M * p.HP; v.qv = t.qv + (size_t)l * M * p.HP;
    v.kb = t.kb + (size_t)l * M * p.HP; v.vb = t.vb + (size_t)l * M * p.HP;
    v.posb = t.posb + (size_t)l * p.R * p.HP;
    return v;
}

// out (M, N) fp32 = x (M, K) W (N, K)^T + bias
int tc_linear(Ctx &c, const bf16 *xb, int ldx, const bf16 *Wb, const float *bias, float *out, int ldc, int M, int N, int K, const char *tag, int relu = 0) {
    BgProblem g;
    g.M = M; g.N = N; g.K = K;
    g.a = {xb, 0, ldx, 0, 0};
    g.b = {Wb, 0, K, 0, 0};
    g.out = out; g.ldc = ldc; g.bias = bias; g.relu = relu;
    return launch_bgemm(c.h, g, c.st, tag);
}
// dx (M, K) fp32 = dy (M, N) W (N, K)
int tc_dgrad(Ctx &c, const bf16 *dyb, int ldy, const bf16 *Wb, float *dx, int ldo, int M, int N, int K) {
    BgProblem g;
    g.M = M; g.N = K; g.K = N;
    g.a = {dyb, 0, ldy, 0, 0};
    g.b = {Wb, 1, K, 0, 0};             // W stored [n_w][k_w]: the contraction index runs along the rows
    g.out = dx; g.ldc = ldo;
    return launch_bgemm(c.h, g, c.st, "train_tc_dgrad");
}
// dW (N, K) fp32 = dy (M, N)^T x (M, K); the contraction over the M frames is split over one wave of CTAs into partial products
// that are summed in slice order; runs on the weight-gradient stream (Ctx::fork / join)
int tc_wgrad(Ctx &c, const bf16 *dyb, int ldy, const bf16 *xb, int ldx, float *dW, int M, int N, int K) {
    BgProblem g;
    g.M = N; g.N = K; g.K = M;
    g.a = {dyb, 1, ldy, 0, 0};
    g.b = {xb, 1, ldx, 0, 0};
    const int bn = K <= 256 ? K : 256;
    const int tiles = ((N + 127) / 128) * ((K + bn - 1) / bn);
    int ks = 148 / (tiles < 1 ? 1 : tiles);            // one wave of CTAs: fewer partial products to sum, longer contraction per CTA
    const int nkc = (M + 63) / 64;
    if (ks > nkc) ks = nkc;
    while (ks > 1 && (size_t)ks * N * K > c.part_floats) --ks;
    if (ks <= 1) {
        g.out = dW; g.ldc = K;
        return launch_bgemm(c.h, g, c.wstream(), "train_tc_wgrad");
    }
    // partial products of the k-slices, summed in slice order (no atomics: bit-reproducible)
    const int kchunks = (nkc + ks - 1) / ks;
    ks = (nkc + kchunks - 1) / kchunks;
    g.out = c.part; g.ldc = K; g.out_ks = (long long)N * K; g.ksplit = ks;
    COCR_TRY(launch_bgemm(c.h, g, c.wstream(), "train_tc_wgrad"));
    k_reduce_z<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, c.wstream()>>>(c.part, ks, (int64_t)N * K, dW);
    TK_LAUNCHED(c);
    return COCR_OK;
}
// out = (x ? x : 0) + factor * dropout(y) [+ bf16 copy] [+ column sums]: the 4-wide kernel when the shape allows, else the scalar ones
int tc_resid_drop(Ctx &c, const float *x, const float *y, float *out, float factor, int64_t rows, int cols, const Drop &dr, bf16 *outb, float *colsum_out) {
    if (cols % 4 == 0) {
        const int rpb = 16;
        k_resid_drop_v4<<<(unsigned)((rows + rpb - 1) / rpb), 256, colsum_out ? (size_t)cols * 4 : 0, c.st>>>(x, y, out, factor, rows, cols, dr, outb, colsum_out, rpb);
        TK_LAUNCHED(c);
        return COCR_OK;
    }
    k_resid_drop<<<ew_grid(rows * cols), 256, 0, c.st>>>(x, y, out, factor, rows * cols, dr, outb);
    TK_LAUNCHED(c);
    if (colsum_out) COCR_TRY(colsum(c, out, cols, (int)rows, cols, colsum_out));
    return COCR_OK;
}
int tc_cast(Ctx &c, const float *src, int64_t lds, bf16 *dst, int64_t ldd, int64_t rows, int cols) {
    if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0))
        k_cast_bf16_v4<<<ew_grid(rows * cols / 4), 256, 0, c.st>>>(src, lds, dst, ldd, rows, cols);
    else
        k_cast_bf16<<<ew_grid(rows * cols), 256, 0, c.st>>>(src, lds, dst, ldd, rows, cols);
    TK_LAUNCHED(c);
    return COCR_OK;
}
// LayerNorm forward straight to the bf16 GEMM operand (+ statistics for the backward pass)
int tc_ln_fwd(Ctx &c, const float *x, const float *g, const float *b, float *stats, int M, int d, bf16 *yb) {
    ProfScope ps(c.h, "train_layernorm", c.st, 0.0, 6.0 * M * d);
    if (d % 4 == 0 && d <= 256) k_ln_fwd_v4<<<(M + 7) / 8, 256, 0, c.st>>>(x, g, b, nullptr, stats, M, d, yb);
    else k_ln_fwd<<<(M + 7) / 8, 256, 0, c.st>>>(x, g, b, nullptr, stats, M, d, yb);
    TK_LAUNCHED(c);
    return COCR_OK;
}

int ffn_forward_tc(Mod &m, int half, const float *x_in, float *x_out) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    const int M = m.p.M, d = dm.d, ff = dm.ff;
    float *y = m.f(m.p.t_d1);
    float *h1 = m.f(m.p.h1) + ((size_t)m.l * 2 + half) * M * ff;
    bf16 *xnb = sv.xnb[half ? 3 : 0], *actb = sv.actb[half];
    COCR_TRY(tc_ln_fwd(m.c, x_in, m.P + o.ffn_ln_g[half], m.P + o.ffn_ln_b[half], m.stats(half ? 3 : 0), M, d, xnb));
    COCR_TRY(tc_linear(m.c, xnb, d, t.pb + o.ffn_w1[half], m.P + o.ffn_b1[half], h1, ff, M, ff, d, "train_tc_ffn_up"));
    if (ff % 4 == 0) k_silu_drop_v4<<<ew_grid((int64_t)M * ff / 4), 256, 0, m.c.st>>>(h1, (int64_t)M * ff, m.drop(m.cfg->feed_forward_dropout_p, half ? 6 : 0), actb);
    else k_silu_drop<<<ew_grid((int64_t)M * ff), 256, 0, m.c.st>>>(h1, nullptr, (int64_t)M * ff, m.drop(m.cfg->feed_forward_dropout_p, half ? 6 : 0), actb);
    TK_LAUNCHED(m.c);
    COCR_TRY(tc_linear(m.c, actb, ff, t.pb + o.ffn_w2[half], m.P + o.ffn_b2[half], y, d, M, d, ff, "train_tc_ffn_down"));
    COCR_TRY(tc_resid_drop(m.c, x_in, y, x_out, dm.ffn_factor, M, d, m.drop(m.cfg->feed_forward_dropout_p, half ? 7 : 1), nullptr, nullptr));
    return COCR_OK;
}

int ffn_backward_tc(Mod &m, int half, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    const int M = m.p.M, d = dm.d, ff = dm.ff;
    float *dy = m.f(m.p.t_d0), *dxn = m.f(m.p.t_d2), *dact = m.f(m.p.t_ff1);
    const float *h1 = m.f(m.p.h1) + ((size_t)m.l * 2 + half) * M * ff;
    const bf16 *xnb = sv.xnb[half ? 3 : 0], *actb = sv.actb[half];
    COCR_TRY(m.c.join());      // the previous module's weight gradients still read dyb / wideb
    // dy = factor * mask * dx, its bf16 copy and the bias gradient of the second linear layer in one pass
    COCR_TRY(tc_resid_drop(m.c, nullptr, dx, dy, dm.ffn_factor, M, d, m.drop(m.cfg->feed_forward_dropout_p, half ? 7 : 1), t.dyb, m.G + o.ffn_b2[half]));
    const Drop d0 = m.drop(m.cfg->feed_forward_dropout_p, half ? 6 : 0);
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.dyb, d, actb, ff, m.G + o.ffn_w2[half], M, d, ff));
    COCR_TRY(tc_dgrad(m.c, t.dyb, d, t.pb + o.ffn_w2[half], dact, ff, M, d, ff));
    if (ff % 4 == 0) {
        const int rpb = 8;
        k_silu_drop_bwd_v4<<<(M + rpb - 1) / rpb, 256, (size_t)ff * 4, m.c.st>>>(h1, dact, M, ff, d0, t.wideb, m.G + o.ffn_b1[half], rpb);
        TK_LAUNCHED(m.c);
    } else {
        k_silu_drop_bwd<<<ew_grid((int64_t)M * ff), 256, 0, m.c.st>>>(h1, dact, (int64_t)M * ff, d0, t.wideb);
        TK_LAUNCHED(m.c);
        COCR_TRY(colsum(m.c, dact, ff, M, ff, m.G + o.ffn_b1[half]));
    }
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.wideb, ff, xnb, d, m.G + o.ffn_w1[half], M, ff, d));
    COCR_TRY(tc_dgrad(m.c, t.wideb, ff, t.pb + o.ffn_w1[half], dxn, d, M, ff, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, m.stats(half ? 3 : 0), m.P + o.ffn_ln_g[half], dx, dx, m.G + o.ffn_ln_g[half], m.G + o.ffn_ln_b[half], M, d));
    return COCR_OK;
}

// one contraction per (line, head): operands are head-pitched (rows, nh*dhp) bf16 matrices or (B*nh, T, Tp|Rp) score blocks
struct AttGeom { int B, T, Tp, R, Rp, nh, dh, dhp, HP, d; };

inline BgOperand head_op(const bf16 *p, int mn, const AttGeom &a, bool per_line = true) {   // rows of line b at b*T, head h at column h*dhp
    return {p, mn, a.HP, per_line ? (long long)a.T * a.HP : 0, a.dhp};
}
inline BgOperand score_op(const bf16 *p, int mn, const AttGeom &a, int ld) {                 // block (b*nh + h) of T rows
    return {p, mn, ld, (long long)a.nh * a.T * ld, (long long)a.T * ld};
}

int att_forward_tc(Mod &m, const float *x_in, float *x_out) {
    const Dims &dm = m.dm;
    TrainState *s = m.c.h->train;
    const TrainLayerOff &o = s->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T, R = m.p.R;
    const AttGeom a{B, T, m.p.Tp, R, m.p.Rp, dm.nh, dm.dh, m.p.dhp, m.p.HP, d};
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    float *y = m.f(m.p.t_d1), *pos = m.f(m.p.pos);                        // qkv / pos in fp32 are scratch here: the bf16 operands are what is kept
    float *qkv = m.f(m.p.qkv), *ctx = m.f(m.p.ctx) + (size_t)m.l * M * d;
    const size_t pblk = (size_t)B * dm.nh * T * a.Tp;
    bf16 *Pl = (bf16 *)(m.ws + m.p.Pb16) + (size_t)m.l * 2 * pblk;
    const bool adrop = m.cfg->attention_dropout_p > 0.0f;
    bf16 *Pd = adrop ? Pl + pblk : Pl;
    COCR_TRY(tc_ln_fwd(m.c, x_in, m.P + o.att_ln_g, m.P + o.att_ln_b, m.stats(1), M, d, sv.xnb[1]));
    COCR_TRY(tc_linear(m.c, sv.xnb[1], d, t.pb + o.w_qkv, m.P + o.b_qkv, qkv, 3 * d, M, 3 * d, d, "train_tc_qkv"));
    COCR_TRY(tc_linear(m.c, t.peb, d, t.pb + o.w_pos, nullptr, pos, d, R, d, d, "train_tc_pos_proj"));
    {
        ProfScope ps(m.c.h, "train_tc_att_prep", m.c.st, 0.0, (double)M * (12.0 * d + 8.0 * a.HP));
        if (dm.dh % 4 == 0) k_att_prep_v4<<<ew_grid((int64_t)M * d / 4), 256, 0, m.c.st>>>(qkv, m.P + o.u_bias, m.P + o.v_bias, sv.qu, sv.qv, sv.kb, sv.vb, M, dm.nh, dm.dh, a.dhp);
        else k_att_prep<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(qkv, m.P + o.u_bias, m.P + o.v_bias, sv.qu, sv.qv, sv.kb, sv.vb, M, dm.nh, dm.dh, a.dhp);
        TK_LAUNCHED(m.c);
        k_head_pitch<<<ew_grid((int64_t)R * d), 256, 0, m.c.st>>>(pos, d, 0, nullptr, sv.posb, R, dm.nh, dm.dh, a.dhp);
        TK_LAUNCHED(m.c);
    }
    {   // content scores (q+u) k^T : (T x T) per (line, head)
        BgProblem g;
        g.M = T; g.N = T; g.K = dm.dh; g.nz1 = B; g.nz2 = dm.nh;
        g.a = head_op(sv.qu, 0, a);
        g.b = head_op(sv.kb, 0, a);
        g.out = t.Sc; g.ldc = a.Tp; g.out_z1 = (long long)dm.nh * T * a.Tp; g.out_z2 = (long long)T * a.Tp;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_content"));
    }
    {   // positional scores (q+v) p^T over the 2T-1 offsets : (T x R) per (line, head); p is shared by the lines
        BgProblem g;
        g.M = T; g.N = R; g.K = dm.dh; g.nz1 = B; g.nz2 = dm.nh;
        g.a = head_op(sv.qv, 0, a);
        g.b = head_op(sv.posb, 0, a, false);
        g.out = t.PS; g.ldc = a.Rp; g.out_z1 = (long long)dm.nh * T * a.Rp; g.out_z2 = (long long)T * a.Rp;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_pos"));
    }
    {
        const int64_t rows = (int64_t)B * dm.nh * T;
        ProfScope ps(m.c.h, "train_tc_att_softmax", m.c.st, 0.0, (double)rows * (4.0 * a.Tp + 4.0 * T + 4.0 * a.Tp));
        k_att_softmax_tc<<<(unsigned)((rows + 7) / 8), 256, 0, m.c.st>>>(t.Sc, t.PS, Pl, Pd, rows, T, a.Tp, a.Rp, 1.0f / sqrtf((float)dm.dh),
                                                                        m.drop(m.cfg->attention_dropout_p, 2));
        TK_LAUNCHED(m.c);
    }
    {   // context = dropout(P) v : A = Pd (T x T), B = v stored [j][c]
        BgProblem g;
        g.M = T; g.N = dm.dh; g.K = T; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(Pd, 0, a, a.Tp);
        g.b = head_op(sv.vb, 1, a);
        g.out = ctx; g.ldc = d; g.out_z1 = (long long)T * d; g.out_z2 = dm.dh;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_pv"));
    }
    COCR_TRY(tc_cast(m.c, ctx, d, sv.ctxb, d, M, d));
    COCR_TRY(tc_linear(m.c, sv.ctxb, d, t.pb + o.w_out, m.P + o.b_out, y, d, M, d, d, "train_tc_att_out"));
    COCR_TRY(tc_resid_drop(m.c, x_in, y, x_out, 1.0f, M, d, m.drop(m.cfg->attention_dropout_p, 3), nullptr, nullptr));
    return COCR_OK;
}

int att_backward_tc(Mod &m, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    TrainState *s = m.c.h->train;
    const TrainLayerOff &o = s->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T, R = m.p.R;
    const AttGeom a{B, T, m.p.Tp, R, m.p.Rp, dm.nh, dm.dh, m.p.dhp, m.p.HP, d};
    float *dy = m.f(m.p.t_d0), *dctx = m.f(m.p.t_d1), *dxn = m.f(m.p.t_d2), *dqkv = m.f(m.p.t_qkv), *dpos = m.f(m.p.dpos);
    float *dqc = m.f(m.p.t_ff0), *dqp = m.f(m.p.t_ff1);
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    const float *ctx = m.f(m.p.ctx) + (size_t)m.l * M * d;
    const size_t pblk = (size_t)B * dm.nh * T * a.Tp;
    const bf16 *Pl = (const bf16 *)(m.ws + m.p.Pb16) + (size_t)m.l * 2 * pblk;
    const bool adrop = m.cfg->attention_dropout_p > 0.0f;
    const bf16 *Pd = adrop ? Pl + pblk : Pl;
    const float inv_sqrt = 1.0f / sqrtf((float)dm.dh);
    COCR_TRY(m.c.join());
    COCR_TRY(tc_resid_drop(m.c, nullptr, dx, dy, 1.0f, M, d, m.drop(m.cfg->attention_dropout_p, 3), t.dyb, m.G + o.b_out));
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.dyb, d, sv.ctxb, d, m.G + o.w_out, M, d, d));
    COCR_TRY(tc_dgrad(m.c, t.dyb, d, t.pb + o.w_out, dctx, d, M, d, d));
    {
        ProfScope ps(m.c.h, "train_tc_att_prep", m.c.st, 0.0, (double)M * (8.0 * d + 2.0 * a.HP));
        if (dm.dh % 4 == 0) k_att_bwd_prep_v4<<<ew_grid((int64_t)M * dm.nh), 256, 0, m.c.st>>>(dctx, ctx, t.dctxb, t.delta, M, dm.nh, dm.dh, a.dhp);
        else k_att_bwd_prep<<<ew_grid((int64_t)M * dm.nh), 256, 0, m.c.st>>>(dctx, ctx, t.dctxb, t.delta, M, dm.nh, dm.dh, a.dhp);
        TK_LAUNCHED(m.c);
    }
    {   // d dropout(P) = dctx v^T : (T x T)
        BgProblem g;
        g.M = T; g.N = T; g.K = dm.dh; g.nz1 = B; g.nz2 = dm.nh;
        g.a = head_op(t.dctxb, 0, a);
        g.b = head_op(sv.vb, 0, a);
        g.out = t.Sc; g.ldc = a.Tp; g.out_z1 = (long long)dm.nh * T * a.Tp; g.out_z2 = (long long)T * a.Tp;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dp"));
    }
    {
        const int64_t rows = (int64_t)B * dm.nh * T;
        ProfScope ps(m.c.h, "train_tc_att_ds", m.c.st, 0.0, (double)rows * (6.0 * a.Tp + 2.0 * a.Tp + 2.0 * a.Rp));
        const size_t ds_smem = (size_t)8 * a.Rp * 2;
        if (T % 4 == 0 && ds_smem <= 48 * 1024)
            k_att_ds_tc_v4<<<(unsigned)((rows + 7) / 8), 256, ds_smem, m.c.st>>>(Pl, t.Sc, t.delta, t.dSb, t.dPSb, rows, T, a.Tp, a.Rp, dm.nh, inv_sqrt,
                                                                                m.drop(m.cfg->attention_dropout_p, 2));
        else
            k_att_ds_tc<<<(unsigned)((rows + 7) / 8), 256, 0, m.c.st>>>(Pl, t.Sc, t.delta, t.dSb, t.dPSb, rows, T, a.Tp, a.Rp, dm.nh, inv_sqrt,
                                                                       m.drop(m.cfg->attention_dropout_p, 2));
        TK_LAUNCHED(m.c);
    }
    {   // dv_j = sum_i dropout(P)[i,j] dctx_i : A = Pd stored [i][j], B = dctx stored [i][c]
        BgProblem g;
        g.M = T; g.N = dm.dh; g.K = T; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(Pd, 1, a, a.Tp);
        g.b = head_op(t.dctxb, 1, a);
        g.out = dqkv + 2 * d; g.ldc = 3 * d; g.out_z1 = (long long)T * 3 * d; g.out_z2 = dm.dh;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dv"));
    }
    {   // dk_j = sum_i dS[i,j] (q_i + u)
        BgProblem g;
        g.M = T; g.N = dm.dh; g.K = T; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(t.dSb, 1, a, a.Tp);
        g.b = head_op(sv.qu, 1, a);
        g.out = dqkv + d; g.ldc = 3 * d; g.out_z1 = (long long)T * 3 * d; g.out_z2 = dm.dh;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dk"));
    }
    {   // content part of dq: sum_j dS[i,j] k_j
        BgProblem g;
        g.M = T; g.N = dm.dh; g.K = T; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(t.dSb, 0, a, a.Tp);
        g.b = head_op(sv.kb, 1, a);
        g.out = dqc; g.ldc = d; g.out_z1 = (long long)T * d; g.out_z2 = dm.dh;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dqc"));
    }
    {   // positional part of dq: sum_r dPS[i,r] p_r
        BgProblem g;
        g.M = T; g.N = dm.dh; g.K = R; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(t.dPSb, 0, a, a.Rp);
        g.b = head_op(sv.posb, 1, a, false);
        g.out = dqp; g.ldc = d; g.out_z1 = (long long)T * d; g.out_z2 = dm.dh;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dqp"));
    }
    COCR_TRY(colsum(m.c, dqc, d, M, d, m.G + o.u_bias));
    COCR_TRY(colsum(m.c, dqp, d, M, d, m.G + o.v_bias));
    k_add2<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(dqc, dqp, dqkv, 3 * d, M, d);
    TK_LAUNCHED(m.c);
    {   // dp_r = sum_b sum_i dPS[b,h][i,r] (q_i + v) : summed over the lines with atomics
        COCR_CUDA(m.c.h, cudaMemsetAsync(dpos, 0, (size_t)R * d * 4, m.c.st));
        BgProblem g;
        g.M = R; g.N = dm.dh; g.K = T; g.nz1 = B; g.nz2 = dm.nh;
        g.a = score_op(t.dPSb, 1, a, a.Rp);
        g.b = head_op(sv.qv, 1, a);
        g.out = dpos; g.ldc = d; g.out_z1 = 0; g.out_z2 = dm.dh; g.out_mode = 2;
        COCR_TRY(launch_bgemm(m.c.h, g, m.c.st, "train_tc_att_dpos"));
    }
    COCR_TRY(tc_cast(m.c, dpos, d, t.dposb, d, R, d));
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.dposb, d, t.peb, d, m.G + o.w_pos, R, d, d));
    COCR_TRY(tc_cast(m.c, dqkv, 3 * d, t.wideb, 3 * d, M, 3 * d));
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.wideb, 3 * d, sv.xnb[1], d, m.G + o.w_qkv, M, 3 * d, d));
    COCR_TRY(colsum(m.c, dqkv, 3 * d, M, 3 * d, m.G + o.b_qkv));
    COCR_TRY(tc_dgrad(m.c, t.wideb, 3 * d, t.pb + o.w_qkv, dxn, d, M, 3 * d, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, m.stats(1), m.P + o.att_ln_g, dx, dx, m.G + o.att_ln_g, m.G + o.att_ln_b, M, d));
    return COCR_OK;
}

int conv_forward_tc(Mod &m, const float *x_in, float *x_out, float *params_rw) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T;
    float *y = m.f(m.p.t_d1), *g = m.f(m.p.t_d2);
    float *u = m.f(m.p.u) + (size_t)m.l * M * 2 * d, *cd = m.f(m.p.cdw) + (size_t)m.l * M * d, *bst = m.f(m.p.bnstat) + (size_t)m.l * 2 * d;
    double *sums = (double *)m.f(m.p.bnsums);
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    COCR_TRY(tc_ln_fwd(m.c, x_in, m.P + o.conv_ln_g, m.P + o.conv_ln_b, m.stats(2), M, d, sv.xnb[2]));
    COCR_TRY(tc_linear(m.c, sv.xnb[2], d, t.pb + o.w_pw1, m.P + o.b_pw1, u, 2 * d, M, 2 * d, d, "train_tc_pw1"));
    if (d % 4 == 0) k_glu_fwd_v4<<<ew_grid((int64_t)M * d / 4), 256, 0, m.c.st>>>(u, g, M, d);
    else k_glu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, g, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(dwconv(m.c, g, m.P + o.w_dw, cd, B, T, d, dm.K, 0));
    COCR_CUDA(m.c.h, cudaMemsetAsync(sums, 0, (size_t)2 * d * 8, m.c.st));
    const int rpb = 32;
    k_bn_sums<<<dim3((M + rpb - 1) / rpb, (d + 255) / 256), 256, 0, m.c.st>>>(cd, M, d, rpb, sums);
    TK_LAUNCHED(m.c);
    k_bn_finalize<<<(d + 255) / 256, 256, 0, m.c.st>>>(sums, M, d, m.cfg->bn_momentum, bst, params_rw ? params_rw + o.bn_rm : nullptr,
                                                       params_rw ? params_rw + o.bn_rv : nullptr);
    TK_LAUNCHED(m.c);
    if (d % 4 == 0) k_bn_silu_fwd_v4<<<ew_grid((int64_t)M * d / 4), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, M, d, sv.zb);
    else k_bn_silu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, nullptr, M, d, sv.zb);
    TK_LAUNCHED(m.c);
    COCR_TRY(tc_linear(m.c, sv.zb, d, t.pb + o.w_pw2, m.P + o.b_pw2, y, d, M, d, d, "train_tc_pw2"));
    COCR_TRY(tc_resid_drop(m.c, x_in, y, x_out, 1.0f, M, d, m.drop(m.cfg->conv_dropout_p, 4), nullptr, nullptr));
    return COCR_OK;
}

int conv_backward_tc(Mod &m, const float *x_in, float *dx) {
    const Dims &dm = m.dm;
    const TrainLayerOff &o = m.c.h->train->L[m.l];
    const TcBuf t = tc_buffers(m.ws, m.p);
    const int M = m.p.M, d = dm.d, B = m.p.B, T = m.p.T;
    float *dy = m.f(m.p.t_d0), *dz = m.f(m.p.t_d2), *du = m.f(m.p.t_ff0), *g = m.f(m.p.t_ff1), *dg = m.f(m.p.t_d1);
    const float *u = m.f(m.p.u) + (size_t)m.l * M * 2 * d, *cd = m.f(m.p.cdw) + (size_t)m.l * M * d, *bst = m.f(m.p.bnstat) + (size_t)m.l * 2 * d;
    double *sums = (double *)m.f(m.p.bnsums);
    COCR_TRY(m.c.join());
    COCR_TRY(tc_resid_drop(m.c, nullptr, dx, dy, 1.0f, M, d, m.drop(m.cfg->conv_dropout_p, 4), t.dyb, m.G + o.b_pw2));
    const TcSaved sv = tc_saved(t, m.p, dm, m.l);
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.dyb, d, sv.zb, d, m.G + o.w_pw2, M, d, d));
    COCR_TRY(tc_dgrad(m.c, t.dyb, d, t.pb + o.w_pw2, dz, d, M, d, d));
    COCR_CUDA(m.c.h, cudaMemsetAsync(sums, 0, (size_t)2 * d * 8, m.c.st));
    const int rpb = 32;
    k_bn_bwd_reduce<<<dim3((M + rpb - 1) / rpb, (d + 255) / 256), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, m.P + o.bn_b, dz, M, d, rpb, sums);
    TK_LAUNCHED(m.c);
    k_bn_bwd_apply<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(cd, bst, m.P + o.bn_g, sums, dz, M, d, m.G + o.bn_g, m.G + o.bn_b);
    TK_LAUNCHED(m.c);
    float *dc = dz;
    if (d % 4 == 0) k_glu_fwd_v4<<<ew_grid((int64_t)M * d / 4), 256, 0, m.c.st>>>(u, g, M, d);
    else k_glu_fwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, g, M, d);
    TK_LAUNCHED(m.c);
    COCR_TRY(dwconv_wgrad(m.c, g, dc, m.G + o.w_dw, B, T, d, dm.K));
    COCR_TRY(dwconv(m.c, dc, m.P + o.w_dw, dg, B, T, d, dm.K, 1));
    if (d % 4 == 0) k_glu_bwd_v4<<<ew_grid((int64_t)M * d / 4), 256, 0, m.c.st>>>(u, dg, du, M, d, t.wideb);
    else k_glu_bwd<<<ew_grid((int64_t)M * d), 256, 0, m.c.st>>>(u, dg, du, M, d, t.wideb);
    TK_LAUNCHED(m.c);
    COCR_TRY(m.c.fork());
    COCR_TRY(tc_wgrad(m.c, t.wideb, 2 * d, sv.xnb[2], d, m.G + o.w_pw1, M, 2 * d, d));
    COCR_TRY(colsum(m.c, du, 2 * d, M, 2 * d, m.G + o.b_pw1));
    float *dxn = dz;
    COCR_TRY(tc_dgrad(m.c, t.wideb, 2 * d, t.pb + o.w_pw1, dxn, d, M, 2 * d, d));
    COCR_TRY(ln_bwd(m.c, dxn, x_in, m.stats(2), m.P + o.conv_ln_g, dx, dx, m.G + o.conv_ln_g, m.G + o.conv_ln_b, M, d));
    return COCR_OK;
}
