// imma_contract.cuh -- the binomial score contraction on the INT8 tensor cores, exact to FP64.
//
// FullStat of the logistic solver (ElasticNetBinaryNEmainEff.c:1693-1735) needs, for every candidate c,
//     G[p][c] = sum_h x_c[h] (w[h] phi_p[h]),   ze[c] = sum_h x_c[h] e[h],   bb[c] = sum_h x_c[h]^2 w[h]
// again after every block of actions, because the IRLS weights w change: 2 N Kc (M + 2) flops, 45 % of a fit's time
// on the FP64 pipe (B200: 37 TFLOP/s, DMMA and DFMA alike).  For genotype-coded designs x is an exact small
// integer, so only the right-hand sides carry 53-bit mantissas.  Each right-hand-side column b is scaled by a power of
// two to |b| <= 1/2 and cut into IM_SLICES signed 7-bit digits (b = sum_s d_s 128^-(s+1), digits in [-64, 64], every
// step exact in FP64); the digit planes are multiplied by the int8 training matrix with mma.sync.m16n8k32.s8 (exact
// int32 accumulation: |sum| <= 64 * 11 * N), and the planes are recombined in FP64 by Horner's rule.  The result equals
// the exact dot product of x_c with b truncated 56 bits below the column's largest entry -- an error bound of
// N * 2^-57 * max|b| * max|x|, below the rounding error of the FP64 summation it replaces -- at a small fraction of its cost.
// Conditions (checked on the host, FoldData::XT8f non-null): main-effect design, every entry an integer, |x| <= 11.
#pragma once
#include "common.cuh"

namespace pareben {

constexpr int IM_NT = 5;          // 8-column tiles per pass of a warp: 20 int32 accumulators + 20 FP64 partial results

__device__ inline void imma_16x8x32(int (&d)[4], const int (&a)[4], int b0, int b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Both operands are stored in FRAGMENT-MAJOR order so that every warp-wide load is one contiguous 512-byte request (the
// natural row-major layouts give eight 16-byte pieces of eight different rows per request and are bound by the L1 tag
// stage, measured 3.5x slower than the FP64 path).  A "k-pair" is 64 consecutive row positions = two k-steps of the
// m16n8k32 instruction; lane (g = lane / 4, tig = lane % 4) owns positions 16 tig .. 16 tig + 15 of it as one int4 whose
// words 0, 1 feed the first instruction (k slots 4 tig.., 16 + 4 tig..) and words 2, 3 the second -- any assignment of
// positions to k slots is valid as long as both operands use the same one.
//   A (training matrix, built once per fold by fragment_major_kernel in pareben.cu):
//       int4 index ((mt * KP + kp) * 32 + lane) * 2 + half       candidate 16 mt + 8 half + g
//   B (digit planes, written by imma_slice):
//       int4 index ((s * NTt + nt) * KP + kp) * 32 + lane        column 8 nt + g of slice s
// bscale[r] = the power of two that undoes column r's scaling.  val(r, h) is the FP64 value of column r at row h
// (positions are the training matrix' permuted row order, see transpose_pad_kernel).  A warp per column, a lane per
// four positions; the digits are peeled off a 56-bit fixed-point integer (round to nearest once, then exact).
template <class Val>
__device__ inline void imma_slice(int R, int Rp, int N, int ldt, Val val, int8_t *__restrict__ bs, double *__restrict__ bscale)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int KP = ldt >> 6, NTt = Rp >> 3;
    for (int r = wid; r < R; r += nw) {
        double mx = 0.0;
        bool bad = false;
        for (int h = lane; h < N; h += 32) { const double v = fabs(val(r, h)); if (!(v <= 1.7e308)) bad = true; mx = fmax(mx, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad = __any_sync(0xffffffffu, bad);
        int ex = 0;
        (void)frexp(mx, &ex);                                   // mx = m 2^ex, m in [1/2, 1)
        const bool live = !bad && mx > 1e-290;                  // an all-zero (or denormal) column contributes zeros
        const double down = live ? scalbn(1.0, 55 - ex) : 0.0;  // |val * down| < 2^55
        if (lane == 0) bscale[r] = bad ? nan("") : (live ? scalbn(1.0, ex + 1) : 0.0);
        const int nt = r >> 3, g = r & 7;
        for (int p4 = 4 * lane; p4 < ldt; p4 += 128) {
            unsigned wout[IM_SLICES];
#pragma unroll
            for (int s = 0; s < IM_SLICES; s++) wout[s] = 0u;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int pos = p4 + b;
                const int h = (pos & ~15) | ((pos & 3) << 2) | ((pos >> 2) & 3);
                long long q = (h < N && live) ? __double2ll_rn(val(r, h) * down) : 0LL;      // value * 2^56 / 2^(ex+1), rounded once
#pragma unroll
                for (int s = IM_SLICES - 1; s >= 0; s--) {      // balanced base-128 digits, least significant first
                    const int d = (int)((q + 64) & 127) - 64;   // in [-64, 63]
                    q = (q - d) >> 7;                           // exact
                    if (s == 0) wout[0] |= ((unsigned)(d + (int)(q << 7)) & 0xffu) << (8 * b);       // top digit takes what is left (|.| <= 64)
                    else wout[s] |= ((unsigned)d & 0xffu) << (8 * b);
                }
            }
            const int kp = p4 >> 6, tig = (p4 >> 4) & 3, word = (p4 >> 2) & 3;
#pragma unroll
            for (int s = 0; s < IM_SLICES; s++)
                *reinterpret_cast<unsigned *>(bs + ((((size_t)s * NTt + nt) * KP + kp) * 32 + g * 4 + tig) * 16 + word * 4) = wout[s];
        }
    }
}

// emit(r, c, value, half) receives  bscale[r] * sum_s 128^-(s+1) * sum_pos A[c][pos] * B[s][r][pos]  for r in [0, R), c in [0, Kc),
// (half = 0 / 1: the lane's lower / upper candidate of the tile), and emit_sq(c, value) the same sum for column 0 against the squared codes A8sqf (the binomial sum_h w x^2: column 0 is
// w o phi_0 = w, the intercept's column of ones).
// The digit planes are what every candidate tile needs again, so they are staged in shared memory: the block walks
// (group of 8 candidate tiles) x (chunk of IM_NT column tiles) x (slice) x (IM_KC k-pairs) stages through a two-deep
// cp.async ring (one 512-byte fragment row per (column tile, k-pair), conflict-free 16-byte reads), a warp per candidate
// tile; the training-matrix fragments come straight from L1/L2.  Reading the planes from L2 once per candidate tile
// instead was measured L2-bandwidth-bound (8.6 MB per contraction and block).
constexpr int IM_KC = 7;          // k-pairs (64 row positions each) per stage: 2 x 5 x 7 x 512 bytes = 35,840 bytes of shared memory

template <class Emit, class EmitSq>
__device__ inline void imma_pass(const int8_t *__restrict__ A8f, const int8_t *__restrict__ A8sqf, int ldt, int Kc, int R, int Rp,
                                 const int8_t *__restrict__ bs, const double *__restrict__ bscale, double *smem, Emit emit, EmitSq emit_sq)
{
    const int T = blockDim.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = T >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int n_mtile = (Kc + 15) >> 4;
    const int KP = ldt >> 6, NTt = Rp >> 3;
    const int nkc = (KP + IM_KC - 1) / IM_KC;
    int4 *sb = reinterpret_cast<int4 *>(smem);
    constexpr int STAGE_INT4 = IM_NT * IM_KC * 32;
    const int4 *__restrict__ bsv = reinterpret_cast<const int4 *>(bs);
    for (int mt0 = 0; mt0 < n_mtile; mt0 += nw) {
        const int mt = mt0 + wid;
        const bool live = mt < n_mtile;
        const int c_lo = mt * 16 + g, c_hi = c_lo + 8;
        const size_t aoff = ((size_t)(live ? mt : 0) * KP * 32 + lane) * 2;
        const int4 *__restrict__ Af = reinterpret_cast<const int4 *>(A8f) + aoff;
        const int4 *__restrict__ Asq = reinterpret_cast<const int4 *>(A8sqf) + aoff;
        for (int r0 = 0; r0 < Rp; r0 += 8 * IM_NT) {
            const int nt_live = min(IM_NT, (Rp - r0) >> 3);
            const bool with_sq = r0 == 0;
            double v[IM_NT][4], vsq[2];
#pragma unroll
            for (int n = 0; n < IM_NT; n++)
#pragma unroll
                for (int i = 0; i < 4; i++) v[n][i] = 0.0;
            vsq[0] = vsq[1] = 0.0;
            int acc[IM_NT][4], asq[4];
            const int n_stage = IM_SLICES * nkc;
            auto issue = [&](int st) {
                const int s = IM_SLICES - 1 - st / nkc, kc0 = (st % nkc) * IM_KC, kcn = min(IM_KC, KP - kc0);
                int4 *dst = sb + (st & 1) * STAGE_INT4;
#ifdef MBI_NO_STAGE
                if (false)
#endif
                for (int idx = threadIdx.x; idx < nt_live * kcn * 32; idx += T) {
                    const int l = idx & 31, q = idx >> 5, kk = q % kcn, n = q / kcn;
                    __pipeline_memcpy_async(dst + (n * IM_KC + kk) * 32 + l, bsv + (((size_t)s * NTt + (r0 >> 3) + n) * KP + kc0 + kk) * 32 + l, 16);
                }
                __pipeline_commit();
            };
            __syncthreads();                                      // the ring is free (previous chunk / phase fully consumed)
            issue(0);
            for (int st = 0; st < n_stage; st++) {
                const int kc0 = (st % nkc) * IM_KC, kcn = min(IM_KC, KP - kc0);
                if (st + 1 < n_stage) { issue(st + 1); __pipeline_wait_prior(1); } else __pipeline_wait_prior(0);
                __syncthreads();                                  // stage st has landed for everyone
                if (kc0 == 0) {
#pragma unroll
                    for (int n = 0; n < IM_NT; n++)
#pragma unroll
                        for (int i = 0; i < 4; i++) acc[n][i] = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) asq[i] = 0;
                }
                if (live) {
                    const int4 *buf = sb + (st & 1) * STAGE_INT4 + lane;
#pragma unroll 2
                    for (int kk = 0; kk < kcn; kk++) {
#ifdef MBI_NO_A
                        const int4 alo = make_int4(lane, kk, 1, 2), ahi = alo;
#else
                        const int4 alo = Af[(kc0 + kk) * 64], ahi = Af[(kc0 + kk) * 64 + 1];
#endif
                        int4 b[IM_NT];
#pragma unroll
                        for (int n = 0; n < IM_NT; n++) b[n] = buf[(n * IM_KC + kk) * 32];          // (tiles past nt_live: stale, never emitted)
                        const int a1[4] = {alo.x, ahi.x, alo.y, ahi.y}, a2[4] = {alo.z, ahi.z, alo.w, ahi.w};
#pragma unroll
                        for (int n = 0; n < IM_NT; n++) if (n < nt_live) imma_16x8x32(acc[n], a1, b[n].x, b[n].y);
#pragma unroll
                        for (int n = 0; n < IM_NT; n++) if (n < nt_live) imma_16x8x32(acc[n], a2, b[n].z, b[n].w);
                        if (with_sq) {
#ifdef MBI_NO_A
                            const int4 qlo = alo, qhi = alo;
#else
                            const int4 qlo = Asq[(kc0 + kk) * 64], qhi = Asq[(kc0 + kk) * 64 + 1];
#endif
                            const int q1[4] = {qlo.x, qhi.x, qlo.y, qhi.y}, q2[4] = {qlo.z, qhi.z, qlo.w, qhi.w};
                            imma_16x8x32(asq, q1, b[0].x, b[0].y);
                            imma_16x8x32(asq, q2, b[0].z, b[0].w);
                        }
                    }
                }
                if (kc0 + kcn == KP) {                            // slice complete: Horner step, 128^-1 exact
#pragma unroll
                    for (int n = 0; n < IM_NT; n++)
#pragma unroll
                        for (int i = 0; i < 4; i++) v[n][i] = (v[n][i] + (double)acc[n][i]) * 0.0078125;
                    vsq[0] = (vsq[0] + (double)asq[0]) * 0.0078125; vsq[1] = (vsq[1] + (double)asq[2]) * 0.0078125;     // column 0: fragment entries 0 (row g) and 2 (row g + 8) of lanes with tig == 0
                }
                __syncthreads();                                  // everyone is done with this buffer before stage st + 2 refills it
            }
            if (live) {
#pragma unroll
                for (int n = 0; n < IM_NT; n++) {
                    if (n < nt_live) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int r = r0 + 8 * n + 2 * tig + (i & 1), c = (i < 2) ? c_lo : c_hi;
                            if (r < R && c < Kc) emit(r, c, v[n][i] * bscale[r], i >> 1);
                        }
                    }
                }
                if (with_sq && tig == 0) {
                    const double b0 = bscale[0];
                    if (c_lo < Kc) emit_sq(c_lo, vsq[0] * b0);
                    if (c_hi < Kc) emit_sq(c_hi, vsq[1] * b0);
                }
            }
        }
    }
    __syncthreads();
}

}  // namespace pareben
