// contract_v7.cuh -- experimental contraction: 4 candidates x 8 right-hand sides per thread, int8 X tiles and
// V tiles streamed with a multi-stage cp.async pipeline (dynamic shared memory).
#pragma once
#include <cuda_pipeline.h>
#include <stdint.h>

namespace v7 {
constexpr int CT = 4, RC = 8, TH = 64, XW = 512, NSTAGE = 3;
constexpr int STAGE_BYTES = TH * XW + TH * RC * 8;          // 32 KB + 4 KB
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES;

__device__ inline double i2d(int v) { return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0; }

// X8: [N][ld8] int8 (ld8 multiple of 16, >= Kc + XW), V: [Np][ldv] doubles zero padded (Np multiple of TH, Rp multiple of RC)
template <class Store>
__device__ inline void contract(const int8_t *__restrict__ X8, int ld8, int N, int Kc, int R, const double *__restrict__ V,
                                int ldv, unsigned char *smem, double *cred, Store store, bool sq_first)
{
    const int T = blockDim.x;                       // 256
    const int nchunk = (R + RC - 1) / RC, ntile = (N + TH - 1) / TH, nstep = nchunk * ntile;
    const int ngroup = (Kc + CT - 1) / CT;
    const int Gp = min(128, (ngroup + 31) & ~31);
    int S = T / Gp; S = S >= 8 ? 8 : (S >= 4 ? 4 : (S >= 2 ? 2 : 1));
    const int rps = TH / S;
    const int grp = threadIdx.x % Gp, split = threadIdx.x / Gp;
    auto xs = [&](int stg) { return reinterpret_cast<int8_t *>(smem + stg * STAGE_BYTES); };
    auto vs = [&](int stg) { return reinterpret_cast<double *>(smem + stg * STAGE_BYTES + TH * XW); };
    for (int c0 = 0; c0 < Kc; c0 += Gp * CT) {
        auto stage = [&](int st) {
            if (st < nstep) {
                const int chunk = st / ntile, tile = st - chunk * ntile, stg = st % NSTAGE;
                int8_t *dx = xs(stg); double *dv = vs(stg);
                const int w16 = Gp * CT / 16;
                for (int idx = threadIdx.x; idx < TH * w16; idx += T) {
                    const int h = idx / w16, q = idx - h * w16;
                    __pipeline_memcpy_async(dx + h * XW + q * 16, X8 + (size_t)min(tile * TH + h, N - 1) * ld8 + c0 + q * 16, 16);
                }
                for (int idx = threadIdx.x; idx < TH * (RC / 2); idx += T) {
                    const int h = idx / (RC / 2), q = idx - h * (RC / 2);
                    __pipeline_memcpy_async(dv + h * RC + 2 * q, V + (size_t)(tile * TH + h) * ldv + chunk * RC + 2 * q, 16);
                }
            }
            __pipeline_commit();                    // always commit so the group count stays uniform
        };
        const int cb = c0 + grp * CT;
        const bool live = cb < Kc && split < S;
        double acc[CT][RC];
#pragma unroll
        for (int a = 0; a < CT; a++)
#pragma unroll
            for (int r = 0; r < RC; r++) acc[a][r] = 0.0;
        __syncthreads();
        for (int s = 0; s < NSTAGE - 1; s++) stage(s);
        for (int st = 0; st < nstep; st++) {
            __pipeline_wait_prior(NSTAGE - 2);      // tile st has landed (for this thread's copies)
            __syncthreads();                        // ... and for everyone's; also: everyone is done with tile st-1
            stage(st + NSTAGE - 1);                 // refill the buffer tile st-1 used
            const int chunk = st / ntile, tile = st - chunk * ntile;
            const int r0 = chunk * RC, h0 = tile * TH, stg = st % NSTAGE;
            if (live) {
                const bool sq = sq_first && r0 == 0;
                const int hb = split * rps, he = min(hb + rps, N - h0);
                const int8_t *xt = xs(stg) + grp * CT;
                const double *buf = vs(stg);
#pragma unroll 2
                for (int hh = hb; hh < he; hh++) {
                    const char4 q = *reinterpret_cast<const char4 *>(xt + hh * XW);
                    const double x[CT] = {i2d(q.x), i2d(q.y), i2d(q.z), i2d(q.w)};
                    const double4 *v4 = reinterpret_cast<const double4 *>(buf + hh * RC);
                    const double4 va = v4[0], vb = v4[1];
#pragma unroll
                    for (int a = 0; a < CT; a++) {
                        acc[a][0] = fma(sq ? x[a] * x[a] : x[a], va.x, acc[a][0]);
                        acc[a][1] = fma(x[a], va.y, acc[a][1]); acc[a][2] = fma(x[a], va.z, acc[a][2]);
                        acc[a][3] = fma(x[a], va.w, acc[a][3]); acc[a][4] = fma(x[a], vb.x, acc[a][4]);
                        acc[a][5] = fma(x[a], vb.y, acc[a][5]); acc[a][6] = fma(x[a], vb.z, acc[a][6]);
                        acc[a][7] = fma(x[a], vb.w, acc[a][7]);
                    }
                }
            }
            if (tile == ntile - 1) {
                if (S == 1) {
                    if (live)
#pragma unroll
                        for (int a = 0; a < CT; a++)
#pragma unroll
                            for (int r = 0; r < RC; r++) { if (r0 + r < R && cb + a < Kc) store(r0 + r, cb + a, acc[a][r]); acc[a][r] = 0.0; }
                } else {
                    if (live)
#pragma unroll
                        for (int a = 0; a < CT; a++)
#pragma unroll
                            for (int r = 0; r < RC; r++) { cred[(((size_t)split * Gp + grp) * CT + a) * RC + r] = acc[a][r]; acc[a][r] = 0.0; }
                    __syncthreads();
                    if (live && split == 0)
#pragma unroll
                        for (int a = 0; a < CT; a++)
#pragma unroll
                            for (int r = 0; r < RC; r++)
                                if (r0 + r < R && cb + a < Kc) {
                                    double z = 0.0;
                                    for (int sp = 0; sp < S; sp++) z += cred[(((size_t)sp * Gp + grp) * CT + a) * RC + r];
                                    store(r0 + r, cb + a, z);
                                }
                    __syncthreads();                // cred is reused by the next chunk
                }
            }
        }
        __pipeline_wait_prior(0);
        __syncthreads();
    }
}
}  // namespace v7
