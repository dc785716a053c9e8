// np_smooth3.cuh -- K4, isotropic specialisation (dx == dy == dz on the level, which holds for every
// level of every benchmark configuration): the tile-resident Gauss-Seidel sweep of np_smooth.cuh
// with the same semantics (tile 64 x 16 x TZ, planes ascending, 4 colours c = (i&1) + 2(j&1) per
// plane, previous-sweep values outside the tile), restructured around three facts:
//
//  (1) With fx == fy == fz = f the six face coefficients of mlndlap_adotx_aa are
//      exactly 0.0 and the other twenty are all F = 3f.  In a plane, colour 1 couples to colour 0
//      and colour 3 to colour 2 only through face neighbours, so {0,1} and {2,3} can be relaxed
//      together: 2 block barriers per plane instead of 5, bit-identical to the 4-step order.
//  (2) L phi = s0 phi + F T with s0 = -4 F G (G = sum of the 8 sigma cells, T = sigma-weighted sum
//      of the 20 neighbours), so the update phi + (rhs - L phi)/s0 is (T - rhs/F) / (4 G): the
//      centre value cancels and the diagonal needs one reciprocal of 4G (MUFU.RCP64H + 5 DFMA).
//  (3) The window of plane k+1 read for the "upper" contribution of plane k is, one iteration later,
//      the previous-sweep window of the plane being relaxed; the sigma layer and the thread's own
//      new values are carried in registers too: 43 shared-memory loads per thread-plane instead of 74.
//
// Staging (cp.async ring of de-interleaved rows), tile shape, thread mapping (2x2 patch per thread)
// and HBM traffic (32 B/node variable sigma, 24 B constant) are those of k_smooth_v2.
#pragma once
#include "np_smooth.cuh"

namespace b200np_dev {

// RES: "resident chunk" variant for small levels (few CTAs, TZ <= SM_RES_TZ): every plane and sigma
// layer of the chunk gets its own shared-memory slot and is requested up front, so the march never
// waits on memory (a 4-slot ring leaves a lone CTA per SM exposed to one DRAM/L2 latency per plane).
constexpr int SM_RES_TZ = 8;
constexpr int SM_RES_DOUBLES = (SM_RES_TZ + 2) * SM_PHI_SLOT + (SM_RES_TZ + 1) * SM_SIG_SLOT;

// DIST: slab-decomposed level with the FillBoundary fused into the sweep (NVLink peer memory, flags of
// np_kernels.cuh K10).  Push protocol: the CTAs of the slab's bottom chunk store plane 0 into the lower neighbour's
// upper ghost slot of the output array right after their FIRST plane (in-loop, from registers), the CTAs of the top
// chunk store the slab's last plane into the upper neighbour's lower ghost slot after their march (remote stores);
// the last CTA of a side raises the neighbour's flag for the next exchange, and the neighbour's next sweep reads its
// LOCAL ghost slots.  Who waits, and when (not in the first sweep of a smooth call, whose input halo was filled by a
// standalone exchange or is zero):
//   * bottom chunk: before its first plane, for the lower neighbour's last plane of the previous sweep -- the one
//     dependency that cannot be hidden (planes ascend), one chunk march + NVLink latency per sweep on small levels;
//   * top chunk: only right before it stages the ghost plane for its LAST plane -- by then the upper neighbour's bottom
//     chunk has long pushed (it does so after one plane), so the top chunk starts with everybody else and never stalls.
// Measured before this split (both waits at kernel start, both pushes after the march): 30 us per sweep on a 128^3 slab
// level against 18 us on one GPU (profiles/r2_8gpu_phase_profile_rank0.txt).
// Every other CTA starts at once, and the boundary chunks are scheduled first (blockIdx.z = 0 is the top chunk, 1 the
// bottom chunk).  A deliberately short top chunk hides more latency on small levels but was measured to cost a V-cycle.
// A flag also tells the neighbour that its previous boundary plane is no longer being read (WAR): the bottom chunk of
// sweep s+1 starts after the lower neighbour's top chunk finished sweep s (its last read of that ghost slot), and the
// top chunk of sweep s+1 cannot finish before the upper neighbour's bottom chunk has begun sweep s+1.
struct HaloFused {
    HaloFlags f;           // my: [4], [5] count bottom / top chunk CTAs that have pushed their plane; f.k: this sweep's exchange
    const double* pin_lo;  // plane -1 of pin: my ghost slot, or the reflection plane at a physical end
    const double* pin_hi;  // plane nzl of pin
    double* out_lo;        // lower neighbour's ghost slot (its plane nzl) of pout, or nullptr
    double* out_hi;        // upper neighbour's ghost slot (its plane -1) of pout, or nullptr
    int tztop;             // planes of the top chunk
    int first;             // 1: first sweep of a smooth call -> the input halo is already in place, do not wait
    int more;              // 1: another fused sweep follows -> raise the neighbours' flags for it
};

// one thread of a bottom (SIDE 0) / top (SIDE 1) chunk CTA, after a CTA barrier that follows the push of
// its boundary plane: count the CTA; the last one of the side raises the neighbour's flag for epoch ep+1
template <int SIDE>
__device__ __forceinline__ void halo_report(const HaloFused& H)
{
    const unsigned long long ep = halo_epoch(H.f);
    unsigned long long* flag = SIDE == 0 ? H.f.lo_flag : H.f.hi_flag;
    if (!H.more || !flag) return;
    __threadfence_system();   // this CTA's remote stores are performed before the count becomes visible
    if (atomicAdd(H.f.my + 4 + SIDE, 1ull) == (unsigned long long)gridDim.x * gridDim.y - 1) {
        H.f.my[4 + SIDE] = 0ull;
        __threadfence_system();
        st_release_sys(flag, ep + 1);
    }
}

// tzarg: the z-chunk height; bit 30 (SM_ZERO_IN) set: pin is known to be zero everywhere (first sweep of a
// pre-smooth, cor = 0) -- the staged planes are filled with zeros instead of being read, and the caller
// skips the memset of pin (8 B/node written + 8 B/node read saved on 1 of 16 sweeps per level and V-cycle)
constexpr int SM_ZERO_IN = 1 << 30;

template <bool VAR, bool FULL, bool RES, bool DIST = false>
__device__ __forceinline__ void smooth_iso_body(const Lev& L, const double* __restrict__ pin, double* __restrict__ pout,
                                                const double* __restrict__ rhs, const int tzarg, double* smem,
                                                const HaloFused* H = nullptr)
{
    const int TZ = tzarg & (SM_ZERO_IN - 1);
    const bool zin = (tzarg & SM_ZERO_IN) != 0;
    double* sphi = smem;
    constexpr int NPS = RES ? SM_RES_TZ + 2 : 4, NSS = RES ? SM_RES_TZ + 1 : 3;
    double* ssig = smem + NPS * SM_PHI_SLOT;
    const unsigned sphi_a = (unsigned)__cvta_generic_to_shared(sphi);
    const unsigned ssig_a = (unsigned)__cvta_generic_to_shared(ssig);

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    // DIST: blockIdx.z = 0 is the top chunk, 1.. the other chunks from the bottom upwards
    const int ztop = DIST ? L.nzl - H->tztop : L.nzl;
    const int kc0 = DIST ? (blockIdx.z == 0 ? ztop : ((int)blockIdx.z - 1) * TZ) : (int)blockIdx.z * TZ;
    const int kc1 = DIST ? (blockIdx.z == 0 ? L.nzl : min(kc0 + TZ, ztop)) : min(kc0 + TZ, L.nzl);
    const bool anyD = lev_any_masked(L);
    // shared-memory slot of node plane p / sigma layer c
    auto pslot = [&](int p) { return RES ? p - kc0 + 1 : (p + 1) & 3; };
    auto sslot = [&](int c) { return RES ? c - kc0 + 1 : (c + 1) % 3; };

    // ---- staging tables (fixed for the whole march): source offset (elements), smem byte offset ----
    int psrc[5], csrc[5];
    unsigned pdst[5], cdst[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int idx = tid + s * 256;
        psrc[s] = -1; csrc[s] = -1; pdst[s] = 0; cdst[s] = 0;
        if (idx < 66 * 18) {
            const int li = idx % 66 - 1, lj = idx / 66 - 1;
            const int gi = i0 + li, gj = j0 + lj;
            pdst[s] = ((lj + 1) * SM_ROW + sm_col(li)) * 8u;
            const bool ok = FULL || ((L.per[0] ? gi <= L.n[0] : gi <= L.n[0] + 1) && (L.per[1] ? gj <= L.n[1] : gj <= L.n[1] + 1));
            if (ok) psrc[s] = nmap(gj, L.n[1], L.per[1]) * L.px + nmap(gi, L.n[0], L.per[0]);
            else for (int q = 0; q < NPS; ++q) sphi[q * SM_PHI_SLOT + (lj + 1) * SM_ROW + sm_col(li)] = 0.0;
        }
        if (VAR && idx < 65 * 17) {
            const int ci = idx % 65 - 1, cj = idx / 65 - 1;
            const int gi = i0 + ci, gj = j0 + cj;
            cdst[s] = ((cj + 1) * SM_ROW + sm_ccol(ci)) * 8u;
            if (FULL || (gi <= L.n[0] && gj <= L.n[1])) csrc[s] = cmap(gj, L.n[1], L.per[1]) * L.cpx + cmap(gi, L.n[0], L.per[0]);
            else for (int q = 0; q < NSS; ++q) ssig[q * SM_SIG_SLOT + (cj + 1) * SM_ROW + sm_ccol(ci)] = 1.0;
        }
    }
    auto issue_phi = [&](int kl) {  // kl in [-1, nzl]
        const double* src = pin + zplane(L, kl) * L.ps;
        if (DIST) { if (kl < 0) src = H->pin_lo; else if (kl >= L.nzl) src = H->pin_hi; }
        unsigned dst = sphi_a + pslot(kl) * (SM_PHI_SLOT * 8);
        asm volatile("" : "+l"(src), "+r"(dst));  // keep the plane base materialised (no per-copy 64-bit multiply)
        if (zin) {   // the slot being refilled was last read before the previous iteration's mid-plane barrier
#pragma unroll
            for (int s = 0; s < 5; ++s)
                if (psrc[s] >= 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst + pdst[s]), "d"(0.0) : "memory");
            return;
        }
#pragma unroll
        for (int s = 0; s < 5; ++s)
            if (psrc[s] >= 0) cp_async8(dst + pdst[s], src + psrc[s]);
    };
    auto issue_sig = [&](int cl) {  // cell layer in [-1, cnzl]
        if (!VAR) return;
        const double* src = L.sigma + czplane(L, cl) * L.cps;
        unsigned dst = ssig_a + sslot(cl) * (SM_SIG_SLOT * 8);
        asm volatile("" : "+l"(src), "+r"(dst));
#pragma unroll
        for (int s = 0; s < 5; ++s)
            if (csrc[s] >= 0) cp_async8(dst + cdst[s], src + csrc[s]);
    };

    // ---- the thread's 2x2 patch ----
    const int gi0 = i0 + 2 * tx, gj0 = j0 + 2 * ty;
    const bool colok[2] = {FULL || gi0 < L.nn[0], FULL || gi0 + 1 < L.nn[0]};
    const bool rowok[2] = {FULL || gj0 < L.nn[1], FULL || gj0 + 1 < L.nn[1]};
    const int roff = gj0 * L.px + gi0;
    auto load_rhs = [&](int kl, double (&r)[2][2], const double* arr = nullptr) {
        const double* q0 = (arr ? arr : rhs) + kl * L.ps + roff;
        asm volatile("" : "+l"(q0));
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double* q = q0 + b * L.px;
            if (FULL) { const double2 v = *reinterpret_cast<const double2*>(q); r[b][0] = v.x; r[b][1] = v.y; }
            else {
                r[b][0] = r[b][1] = 0.0;
                if (rowok[b]) {
                    if (colok[1]) { const double2 v = *reinterpret_cast<const double2*>(q); r[b][0] = v.x; r[b][1] = v.y; }
                    else if (colok[0]) r[b][0] = q[0];
                }
            }
        }
    };

    const double Finv = 1.0 / L.fxyz;                       // F = 3 dxinv^2 / 36
    const double kr = VAR ? 0.0 : Finv / (32.0 * L.csig);   // constant sigma: phi = (C + 2E)/32 - rhs/(32 F sigma)

    // window columns of this thread in a de-interleaved row: li = 2tx-1, 2tx, 2tx+1, 2tx+2
    const int wc[4] = {33 + tx, tx, 34 + tx, tx + 1};
    const int cc[3] = {32 + tx, tx, 33 + tx};  // cells 2tx-1, 2tx, 2tx+1
    const int rbase = (2 * ty) * SM_ROW;       // window row 0 of this thread

    if (pdl_small_grid()) pdl_trigger();
    pdl_wait();   // everything above only touched kernel parameters and shared memory
    // DIST: the bottom chunk needs the lower neighbour's plane (my ghost slot -1) for its very first plane: wait now.
    // The top chunk needs the upper neighbour's plane (ghost slot nzl) only for its LAST plane: that wait is deferred
    // to the moment the ghost plane is staged (wait_hi below), so the top chunk starts with everybody else and the flag --
    // raised by the neighbour's bottom chunk right after ITS first plane -- has long arrived by then.
    const bool dist_top = DIST && kc1 == L.nzl && H->f.hi_flag != nullptr && !H->first;
    if (DIST && !H->first && kc0 == 0 && H->f.lo_flag) {
        if (tid == 0) {
            const unsigned long long ep = halo_epoch(H->f);
            halo_spin(H->f.my, 0, ep);
        }
        __syncthreads();
    }
    auto wait_hi = [&]() {   // called by all threads right before the ghost plane nzl is staged
        if (tid == 0) {
            const unsigned long long ep = halo_epoch(H->f);
            halo_spin(H->f.my, 1, ep);
        }
        __syncthreads();
    };
    bool hi_waited = false;
    double rcur[2][2], rnext[2][2];
    if (RES) {
        // the whole chunk: planes kc0-1 .. kc1, sigma layers kc0-1 .. kc1-1 (the upper ghost plane of a slab's top
        // chunk is staged later, see the march)
        if (dist_top && kc0 >= L.nzl - 1) { wait_hi(); hi_waited = true; }   // one-plane top chunk: no later point to wait at
        for (int p = kc0 - 1; p <= kc1 - ((dist_top && !hi_waited) ? 1 : 0); ++p) issue_phi(p);
        for (int c = kc0 - 1; c < kc1; ++c) issue_sig(c);
        cp_async_commit();
        load_rhs(kc0, rcur);
        cp_async_wait<0>();
    } else {
        // prologue: planes kc0-1, kc0 (+ sigma layer kc0-1), then plane kc0+1 (+ sigma layer kc0)
        issue_phi(kc0 - 1); issue_phi(kc0); issue_sig(kc0 - 1);
        cp_async_commit();
        if (dist_top && kc0 + 2 >= L.nzl) { wait_hi(); hi_waited = true; }   // one- or two-plane top chunk: the ghost plane is staged right away
        issue_phi(kc0 + 1); issue_sig(kc0);
        cp_async_commit();
        load_rhs(kc0, rcur);
        cp_async_wait<1>();
    }
    __syncthreads();

    // state carried from plane to plane
    double SL[3][3];     // sigma layer kl-1
    double ownp[2][2];   // this sweep's values of the patch on plane kl-1 (previous sweep's for kl = kc0)
    double Wn[2][4];     // previous-sweep values of plane kl, window rows 0 and 2
    {
        const double* Pm = sphi + pslot(kc0 - 1) * SM_PHI_SLOT + rbase;
        const double* P0 = sphi + pslot(kc0) * SM_PHI_SLOT + rbase;
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) ownp[b][a] = Pm[(b + 1) * SM_ROW + wc[a + 1]];
#pragma unroll
        for (int c = 0; c < 4; ++c) { Wn[0][c] = P0[wc[c]]; Wn[1][c] = P0[2 * SM_ROW + wc[c]]; }
        if (VAR) {
            const double* S = ssig + sslot(kc0 - 1) * SM_SIG_SLOT + rbase;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) SL[r][c] = S[r * SM_ROW + cc[c]];
        }
    }

    // The march runs in up to three segments so that the slab protocol stays OUT of the plane loop (a barrier or a fence
    // inside it, even when never executed, pins the instruction schedule: measured 152 instead of 119 us per 256^3-slab
    // sweep): [first plane] -> the bottom chunk pushes plane 0 down -> [...] -> the top chunk waits for the upper
    // neighbour and stages the ghost plane -> [the planes that need it].  One GPU: a single segment.
    const bool botpush = DIST && kc0 == 0 && H->out_lo != nullptr;
    const int khi = RES ? L.nzl - 1 : L.nzl - 2;   // first plane whose iteration needs the upper ghost plane staged
    int kl = kc0;
#pragma unroll 1
    for (int seg = DIST ? 0 : 2; seg < 3; ++seg) {
    int kend = kc1;
    if (DIST && seg == 0) kend = botpush ? min(kc0 + 1, kc1) : kc0;
    if (DIST && seg == 1) kend = (dist_top && !hi_waited) ? min(max(kl, khi), kc1) : kl;
#pragma unroll 1
    for (; kl < kend; ++kl) {
        if (!RES) {
            if (kl + 2 <= kc1) { issue_phi(kl + 2); issue_sig(kl + 1); }
            cp_async_commit();
        }
        load_rhs(kl + 1 < kc1 ? kl + 1 : kl, rnext);
        if (!RES) cp_async_wait<1>();   // plane kl+1 / sigma layer kl have landed (this thread's copies)
        __syncthreads();      // ... everybody else's, and plane kl-1's colours 2,3 are published
        const int kg = kl + L.k0;
        const double* Pm = sphi + pslot(kl - 1) * SM_PHI_SLOT + rbase;   // plane kl-1 (this sweep)
        double* P0 = sphi + pslot(kl) * SM_PHI_SLOT + rbase;             // plane kl
        const double* Pp = sphi + pslot(kl + 1) * SM_PHI_SLOT + rbase;   // plane kl+1 (previous sweep)

        double T[2][2];      // sigma-weighted neighbour sums
        double Sz[3][3];     // SL + SU: weights of the in-plane diagonal neighbours
        double rinv[2][2];   // 1 / (4 G)
        double Wp02[2][4];   // rows 0 and 2 of plane kl+1 (next iteration's Wn)
        if (VAR) {
            // ---- lower side: plane kl-1, sigma layer kl-1 ----
            {
                double W[4][4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { W[0][c] = Pm[wc[c]]; W[3][c] = Pm[3 * SM_ROW + wc[c]]; }
#pragma unroll
                for (int r = 1; r < 3; ++r) {
                    W[r][0] = Pm[r * SM_ROW + wc[0]]; W[r][3] = Pm[r * SM_ROW + wc[3]];
                    W[r][1] = ownp[r - 1][0]; W[r][2] = ownp[r - 1][1];
                }
                double hx[3][2], hy[2][3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { hx[r][0] = SL[r][0] + SL[r][1]; hx[r][1] = SL[r][1] + SL[r][2]; }
#pragma unroll
                for (int c = 0; c < 3; ++c) { hy[0][c] = SL[0][c] + SL[1][c]; hy[1][c] = SL[1][c] + SL[2][c]; }
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        double t = SL[b][a] * W[b][a];
                        t = fma(SL[b][a + 1], W[b][a + 2], t);
                        t = fma(SL[b + 1][a], W[b + 2][a], t);
                        t = fma(SL[b + 1][a + 1], W[b + 2][a + 2], t);
                        double u = hx[b][a] * W[b][a + 1];
                        u = fma(hx[b + 1][a], W[b + 2][a + 1], u);
                        u = fma(hy[b][a], W[b + 1][a], u);
                        u = fma(hy[b][a + 1], W[b + 1][a + 2], u);
                        T[b][a] = t + u;
                    }
            }
            // ---- upper side: plane kl+1, sigma layer kl ----
            {
                const double* S = ssig + sslot(kl) * SM_SIG_SLOT + rbase;
                double SU[3][3], W[4][4];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) SU[r][c] = S[r * SM_ROW + cc[c]];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) W[r][c] = Pp[r * SM_ROW + wc[c]];
                double hx[3][2], hy[2][3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { hx[r][0] = SU[r][0] + SU[r][1]; hx[r][1] = SU[r][1] + SU[r][2]; }
#pragma unroll
                for (int c = 0; c < 3; ++c) { hy[0][c] = SU[0][c] + SU[1][c]; hy[1][c] = SU[1][c] + SU[2][c]; }
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        double t = SU[b][a] * W[b][a];
                        t = fma(SU[b][a + 1], W[b][a + 2], t);
                        t = fma(SU[b + 1][a], W[b + 2][a], t);
                        t = fma(SU[b + 1][a + 1], W[b + 2][a + 2], t);
                        double u = hx[b][a] * W[b][a + 1];
                        u = fma(hx[b + 1][a], W[b + 2][a + 1], u);
                        u = fma(hy[b][a], W[b + 1][a], u);
                        u = fma(hy[b][a + 1], W[b + 1][a + 2], u);
                        T[b][a] += t + u;
                    }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) { Sz[r][c] = SL[r][c] + SU[r][c]; SL[r][c] = SU[r][c]; }
#pragma unroll
                for (int c = 0; c < 4; ++c) { Wp02[0][c] = W[0][c]; Wp02[1][c] = W[2][c]; }
            }
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a)
                    rinv[b][a] = rcp_fast(4.0 * ((Sz[b][a] + Sz[b][a + 1]) + (Sz[b + 1][a] + Sz[b + 1][a + 1])));
        } else {
            // constant sigma: T = C + 2E over the summed window of planes kl-1 and kl+1
            double W[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double up = Pp[r * SM_ROW + wc[c]];
                    const bool mine = (r == 1 || r == 2) && (c == 1 || c == 2);
                    const double lo = mine ? ownp[(r - 1) & 1][(c - 1) & 1] : Pm[r * SM_ROW + wc[c]];
                    W[r][c] = lo + up;
                    if (r == 0) Wp02[0][c] = up;
                    if (r == 2) Wp02[1][c] = up;
                }
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const double C = (W[b][a] + W[b][a + 2]) + (W[b + 2][a] + W[b + 2][a + 2]);
                    const double E = (W[b][a + 1] + W[b + 2][a + 1]) + (W[b + 1][a] + W[b + 1][a + 2]);
                    T[b][a] = fma(2.0, E, C);
                }
        }

        double v[2][2];
        // ---- step 1: colours 0 and 1 (patch row 0); their in-plane diagonal neighbours (window rows 0
        //      and 2 of plane kl) are colours 2/3 = previous-sweep values, already in registers ----
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            double t;
            if (VAR) {
                t = fma(Sz[0][a], Wn[0][a], T[0][a]);
                t = fma(Sz[0][a + 1], Wn[0][a + 2], t);
                t = fma(Sz[1][a], Wn[1][a], t);
                t = fma(Sz[1][a + 1], Wn[1][a + 2], t);
                v[0][a] = fma(-rcur[0][a], Finv, t) * rinv[0][a];
            } else {
                t = fma(2.0, (Wn[0][a] + Wn[0][a + 2]) + (Wn[1][a] + Wn[1][a + 2]), T[0][a]);
                v[0][a] = fma(-rcur[0][a], kr, t * 0.03125);
            }
            if (anyD && node_masked(L, gi0 + a, gj0, kg)) v[0][a] = 0.0;
            // a patch node outside the domain holds the staged wrap / reflection image of a real
            // node (previous-sweep value, like any other halo entry): it must not be relaxed
            if (FULL || (colok[a] && rowok[0])) P0[SM_ROW + wc[a + 1]] = v[0][a];
            else v[0][a] = P0[SM_ROW + wc[a + 1]];
        }
        __syncthreads();
        // ---- step 2: colours 2 and 3 (patch row 1); diagonal neighbours are colours 1/0 of window rows
        //      1 and 3, relaxed in step 1 (by this thread, its neighbours, or halo = previous sweep) ----
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double n00 = a ? v[0][0] : P0[SM_ROW + wc[0]];         // (1, a)
            const double n01 = a ? P0[SM_ROW + wc[3]] : v[0][1];         // (1, a+2)
            const double n10 = P0[3 * SM_ROW + wc[a]];                   // (3, a)
            const double n11 = P0[3 * SM_ROW + wc[a + 2]];               // (3, a+2)
            double t;
            if (VAR) {
                t = fma(Sz[1][a], n00, T[1][a]);
                t = fma(Sz[1][a + 1], n01, t);
                t = fma(Sz[2][a], n10, t);
                t = fma(Sz[2][a + 1], n11, t);
                v[1][a] = fma(-rcur[1][a], Finv, t) * rinv[1][a];
            } else {
                t = fma(2.0, (n00 + n01) + (n10 + n11), T[1][a]);
                v[1][a] = fma(-rcur[1][a], kr, t * 0.03125);
            }
            if (anyD && node_masked(L, gi0 + a, gj0 + 1, kg)) v[1][a] = 0.0;
            if (FULL || (colok[a] && rowok[1])) P0[2 * SM_ROW + wc[a + 1]] = v[1][a];
            else v[1][a] = P0[2 * SM_ROW + wc[a + 1]];
        }
        // ---- store the finished plane ----
        {
            double* q0 = pout + kl * L.ps + roff;
            asm volatile("" : "+l"(q0));
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                double* q = q0 + b * L.px;
                if (FULL) *reinterpret_cast<double2*>(q) = make_double2(v[b][0], v[b][1]);
                else if (rowok[b]) {
                    if (colok[1]) *reinterpret_cast<double2*>(q) = make_double2(v[b][0], v[b][1]);
                    else if (colok[0]) q[0] = v[b][0];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) { rcur[b][a] = rnext[b][a]; ownp[b][a] = v[b][a]; }
#pragma unroll
        for (int c = 0; c < 4; ++c) { Wn[0][c] = Wp02[0][c]; Wn[1][c] = Wp02[1][c]; }
        if (kl + 2 == kc1) pdl_trigger();   // tail of the march (no-op if already triggered)
    }
    if (DIST) {
        // ownp: this thread's values of the last finished plane
        auto push = [&](double* r0) {
            r0 += roff;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                double* q = r0 + b * L.px;
                if (FULL) *reinterpret_cast<double2*>(q) = make_double2(ownp[b][0], ownp[b][1]);
                else if (rowok[b]) {
                    if (colok[1]) *reinterpret_cast<double2*>(q) = make_double2(ownp[b][0], ownp[b][1]);
                    else if (colok[0]) q[0] = ownp[b][0];
                }
            }
        };
        if (seg == 0 && botpush && kl == kc0 + 1) {
            // the slab's first plane goes to the lower neighbour's upper ghost slot of pout as soon as it exists (that
            // neighbour's top chunk needs it for its last plane only, but the flag must be there by then)
            push(H->out_lo);
            __syncthreads();   // every thread's remote stores have been issued
            if (tid == 0) halo_report<0>(*H);
        }
        if (seg == 1 && dist_top && !hi_waited && kl == khi && kl < kc1) {
            wait_hi();
            hi_waited = true;
            if (RES) {   // resident chunk: the ghost plane above the slab is staged now, before the last plane
                issue_phi(L.nzl);
                cp_async_commit();
                cp_async_wait<0>();
            }
        }
        if (seg == 2 && kc1 == L.nzl && H->out_hi) {
            // the slab's last plane goes into the upper neighbour's lower ghost slot of pout after the march
            push(H->out_hi);
            __syncthreads();
            if (tid == 0) halo_report<1>(*H);
        }
    }
    }
    pdl_trigger();
    cp_async_wait<0>();
}

// slab-decomposed variants with the halo exchange fused in (see HaloFused)
template <bool VAR>
__global__ void __launch_bounds__(256, 2) k_smooth_iso_dist(const Lev L, const double* __restrict__ pin, double* __restrict__ pout,
                                                            const double* __restrict__ rhs, int TZ, const HaloFused H)
{
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const bool full = (i0 + SM_TX <= L.nn[0]) && (j0 + SM_TY <= L.nn[1]);
    if (full) smooth_iso_body<VAR, true, false, true>(L, pin, pout, rhs, TZ, smem, &H);
    else      smooth_iso_body<VAR, false, false, true>(L, pin, pout, rhs, TZ, smem, &H);
}
template <bool VAR>
__global__ void __launch_bounds__(256, 1) k_smooth_iso_res_dist(const Lev L, const double* __restrict__ pin, double* __restrict__ pout,
                                                                const double* __restrict__ rhs, int TZ, const HaloFused H)
{
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const bool full = (i0 + SM_TX <= L.nn[0]) && (j0 + SM_TY <= L.nn[1]);
    if (full) smooth_iso_body<VAR, true, true, true>(L, pin, pout, rhs, TZ, smem, &H);
    else      smooth_iso_body<VAR, false, true, true>(L, pin, pout, rhs, TZ, smem, &H);
}

template <bool VAR>
__global__ void __launch_bounds__(256, 2) k_smooth_iso(const Lev L, const double* __restrict__ pin,
                                                       double* __restrict__ pout, const double* __restrict__ rhs, int TZ)
{
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const bool full = (i0 + SM_TX <= L.nn[0]) && (j0 + SM_TY <= L.nn[1]);
    if (full) smooth_iso_body<VAR, true, false>(L, pin, pout, rhs, TZ, smem);
    else      smooth_iso_body<VAR, false, false>(L, pin, pout, rhs, TZ, smem);
}

// small levels: one CTA per SM, whole chunk resident (TZ <= SM_RES_TZ)
template <bool VAR>
__global__ void __launch_bounds__(256, 1) k_smooth_iso_res(const Lev L, const double* __restrict__ pin,
                                                           double* __restrict__ pout, const double* __restrict__ rhs, int TZ)
{
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const bool full = (i0 + SM_TX <= L.nn[0]) && (j0 + SM_TY <= L.nn[1]);
    if (full) smooth_iso_body<VAR, true, true>(L, pin, pout, rhs, TZ, smem);
    else      smooth_iso_body<VAR, false, true>(L, pin, pout, rhs, TZ, smem);
}

// ------------------------------------------------------------------------------------------
// K3, isotropic specialisation: res = rhs - L phi with the same cancelled-coefficient form
// (L phi = F (T - 4 G phi)); the window of the plane itself is carried in registers from the
// previous iteration's "upper" window.  One barrier per plane.  Optional per-CTA inf-norm partial.
// ------------------------------------------------------------------------------------------
template <bool VAR, bool FULL>
__device__ __forceinline__ double resid_iso_body(const Lev& L, const double* __restrict__ phi, const double* __restrict__ rhs,
                                                 double* __restrict__ res, const int TZ, double* smem)
{
    double* sphi = smem;
    double* ssig = smem + 4 * SM_PHI_SLOT;
    const unsigned sphi_a = (unsigned)__cvta_generic_to_shared(sphi);
    const unsigned ssig_a = (unsigned)__cvta_generic_to_shared(ssig);
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const int kc0 = blockIdx.z * TZ, kc1 = min(kc0 + TZ, L.nzl);
    const bool anyD = lev_any_masked(L);

    int psrc[5], csrc[5];
    unsigned pdst[5], cdst[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int idx = tid + s * 256;
        psrc[s] = -1; csrc[s] = -1; pdst[s] = 0; cdst[s] = 0;
        if (idx < 66 * 18) {
            const int li = idx % 66 - 1, lj = idx / 66 - 1;
            const int gi = i0 + li, gj = j0 + lj;
            pdst[s] = ((lj + 1) * SM_ROW + sm_col(li)) * 8u;
            const bool ok = FULL || ((L.per[0] ? gi <= L.n[0] : gi <= L.n[0] + 1) && (L.per[1] ? gj <= L.n[1] : gj <= L.n[1] + 1));
            if (ok) psrc[s] = nmap(gj, L.n[1], L.per[1]) * L.px + nmap(gi, L.n[0], L.per[0]);
            else for (int q = 0; q < 4; ++q) sphi[q * SM_PHI_SLOT + (lj + 1) * SM_ROW + sm_col(li)] = 0.0;
        }
        if (VAR && idx < 65 * 17) {
            const int ci = idx % 65 - 1, cj = idx / 65 - 1;
            const int gi = i0 + ci, gj = j0 + cj;
            cdst[s] = ((cj + 1) * SM_ROW + sm_ccol(ci)) * 8u;
            if (FULL || (gi <= L.n[0] && gj <= L.n[1])) csrc[s] = cmap(gj, L.n[1], L.per[1]) * L.cpx + cmap(gi, L.n[0], L.per[0]);
            else for (int q = 0; q < 3; ++q) ssig[q * SM_SIG_SLOT + (cj + 1) * SM_ROW + sm_ccol(ci)] = 1.0;
        }
    }
    auto issue_phi = [&](int kl) {
        const double* src = phi + zplane(L, kl) * L.ps;
        unsigned dst = sphi_a + ((kl + 1) & 3) * (SM_PHI_SLOT * 8);
        asm volatile("" : "+l"(src), "+r"(dst));
#pragma unroll
        for (int s = 0; s < 5; ++s)
            if (psrc[s] >= 0) cp_async8(dst + pdst[s], src + psrc[s]);
    };
    auto issue_sig = [&](int cl) {
        if (!VAR) return;
        const double* src = L.sigma + czplane(L, cl) * L.cps;
        unsigned dst = ssig_a + ((cl + 1) % 3) * (SM_SIG_SLOT * 8);
        asm volatile("" : "+l"(src), "+r"(dst));
#pragma unroll
        for (int s = 0; s < 5; ++s)
            if (csrc[s] >= 0) cp_async8(dst + cdst[s], src + csrc[s]);
    };
    const int gi0 = i0 + 2 * tx, gj0 = j0 + 2 * ty;
    const bool colok[2] = {FULL || gi0 < L.nn[0], FULL || gi0 + 1 < L.nn[0]};
    const bool rowok[2] = {FULL || gj0 < L.nn[1], FULL || gj0 + 1 < L.nn[1]};
    const int roff = gj0 * L.px + gi0;
    auto load_rhs = [&](int kl, double (&r)[2][2]) {
        const double* q0 = rhs + kl * L.ps + roff;
        asm volatile("" : "+l"(q0));
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double* q = q0 + b * L.px;
            if (FULL) { const double2 v = *reinterpret_cast<const double2*>(q); r[b][0] = v.x; r[b][1] = v.y; }
            else {
                r[b][0] = r[b][1] = 0.0;
                if (rowok[b]) {
                    if (colok[1]) { const double2 v = *reinterpret_cast<const double2*>(q); r[b][0] = v.x; r[b][1] = v.y; }
                    else if (colok[0]) r[b][0] = q[0];
                }
            }
        }
    };
    const double F = L.fxyz;
    const double sF = L.csig * F;
    const int wc[4] = {33 + tx, tx, 34 + tx, tx + 1};
    const int cc[3] = {32 + tx, tx, 33 + tx};
    const int rbase = (2 * ty) * SM_ROW;

    if (pdl_small_grid()) pdl_trigger();
    pdl_wait();
    issue_phi(kc0 - 1); issue_phi(kc0); issue_phi(kc0 + 1); issue_sig(kc0 - 1); issue_sig(kc0);
    cp_async_commit();
    double rcur[2][2], rnext[2][2];
    load_rhs(kc0, rcur);
    cp_async_wait<0>();
    __syncthreads();
    double W0[4][4];   // window of plane kl
    double SL[3][3];   // sigma layer kl-1
    {
        const double* P0 = sphi + ((kc0 + 1) & 3) * SM_PHI_SLOT + rbase;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) W0[r][c] = P0[r * SM_ROW + wc[c]];
        if (VAR) {
            const double* S = ssig + ((kc0) % 3) * SM_SIG_SLOT + rbase;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) SL[r][c] = S[r * SM_ROW + cc[c]];
        }
    }
    double amax = 0.0;

#pragma unroll 1
    for (int kl = kc0; kl < kc1; ++kl) {
        if (kl > kc0) {
            cp_async_wait<0>();
            __syncthreads();   // planes kl-1, kl+1 visible; everybody is done with plane kl-2's slot
        }
        if (kl + 2 <= kc1) { issue_phi(kl + 2); issue_sig(kl + 1); }
        cp_async_commit();
        load_rhs(kl + 1 < kc1 ? kl + 1 : kl, rnext);
        const int kg = kl + L.k0;
        const double* Pm = sphi + ((kl) & 3) * SM_PHI_SLOT + rbase;
        const double* Pp = sphi + ((kl + 2) & 3) * SM_PHI_SLOT + rbase;
        double out[2][2];
        double Wp[4][4];
        if (VAR) {
            // acc = 4 G phi - T, built in three stages ordered to keep few registers live:
            // in-plane diagonals (retires W0), lower side (retires SL), upper side (its window is the next W0)
            double acc[2][2];
            double SU[3][3];
            {
                const double* S = ssig + ((kl + 1) % 3) * SM_SIG_SLOT + rbase;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) SU[r][c] = S[r * SM_ROW + cc[c]];
                double Sz[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Sz[r][c] = SL[r][c] + SU[r][c];
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        double t = Sz[b][a] * W0[b][a];
                        t = fma(Sz[b][a + 1], W0[b][a + 2], t);
                        t = fma(Sz[b + 1][a], W0[b + 2][a], t);
                        t = fma(Sz[b + 1][a + 1], W0[b + 2][a + 2], t);
                        const double G4 = 4.0 * ((Sz[b][a] + Sz[b][a + 1]) + (Sz[b + 1][a] + Sz[b + 1][a + 1]));
                        acc[b][a] = fma(G4, W0[b + 1][a + 1], -t);
                    }
            }
            {
                double W[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) W[r][c] = Pm[r * SM_ROW + wc[c]];
                double hx[3][2], hy[2][3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { hx[r][0] = SL[r][0] + SL[r][1]; hx[r][1] = SL[r][1] + SL[r][2]; }
#pragma unroll
                for (int c = 0; c < 3; ++c) { hy[0][c] = SL[0][c] + SL[1][c]; hy[1][c] = SL[1][c] + SL[2][c]; }
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        double t = SL[b][a] * W[b][a];
                        t = fma(SL[b][a + 1], W[b][a + 2], t);
                        t = fma(SL[b + 1][a], W[b + 2][a], t);
                        t = fma(SL[b + 1][a + 1], W[b + 2][a + 2], t);
                        double u = hx[b][a] * W[b][a + 1];
                        u = fma(hx[b + 1][a], W[b + 2][a + 1], u);
                        u = fma(hy[b][a], W[b + 1][a], u);
                        u = fma(hy[b][a + 1], W[b + 1][a + 2], u);
                        acc[b][a] -= t + u;
                    }
            }
            {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) Wp[r][c] = Pp[r * SM_ROW + wc[c]];
                double hx[3][2], hy[2][3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { hx[r][0] = SU[r][0] + SU[r][1]; hx[r][1] = SU[r][1] + SU[r][2]; }
#pragma unroll
                for (int c = 0; c < 3; ++c) { hy[0][c] = SU[0][c] + SU[1][c]; hy[1][c] = SU[1][c] + SU[2][c]; }
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        double t = SU[b][a] * Wp[b][a];
                        t = fma(SU[b][a + 1], Wp[b][a + 2], t);
                        t = fma(SU[b + 1][a], Wp[b + 2][a], t);
                        t = fma(SU[b + 1][a + 1], Wp[b + 2][a + 2], t);
                        double u = hx[b][a] * Wp[b][a + 1];
                        u = fma(hx[b + 1][a], Wp[b + 2][a + 1], u);
                        u = fma(hy[b][a], Wp[b + 1][a], u);
                        u = fma(hy[b][a + 1], Wp[b + 1][a + 2], u);
                        out[b][a] = fma(F, acc[b][a] - (t + u), rcur[b][a]);   // rhs - F (T - 4 G phi)
                    }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) SL[r][c] = SU[r][c];
            }
        } else {
            double W[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) { Wp[r][c] = Pp[r * SM_ROW + wc[c]]; W[r][c] = Pm[r * SM_ROW + wc[c]] + Wp[r][c]; }
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const double C = (W[b][a] + W[b][a + 2]) + (W[b + 2][a] + W[b + 2][a + 2]);
                    const double E = ((W[b][a + 1] + W[b + 2][a + 1]) + (W[b + 1][a] + W[b + 1][a + 2])) +
                                     ((W0[b][a] + W0[b][a + 2]) + (W0[b + 2][a] + W0[b + 2][a + 2]));
                    const double t = fma(2.0, E, C);
                    out[b][a] = fma(sF, fma(32.0, W0[b + 1][a + 1], -t), rcur[b][a]);     // rhs - sigma F (C + 2E - 32 phi)
                }
        }
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if (anyD && node_masked(L, gi0 + a, gj0 + b, kg)) out[b][a] = 0.0;
                if (FULL || (colok[a] && rowok[b])) amax = fmax(amax, fabs(out[b][a]));
            }
        if (res) {   // nullptr: the caller only wants the norm
            double* q0 = res + kl * L.ps + roff;
            asm volatile("" : "+l"(q0));
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                double* q = q0 + b * L.px;
                if (FULL) *reinterpret_cast<double2*>(q) = make_double2(out[b][0], out[b][1]);
                else if (rowok[b]) {
                    if (colok[1]) *reinterpret_cast<double2*>(q) = make_double2(out[b][0], out[b][1]);
                    else if (colok[0]) q[0] = out[b][0];
                }
            }
        }
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a) rcur[b][a] = rnext[b][a];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) W0[r][c] = Wp[r][c];
    }
    pdl_trigger();
    cp_async_wait<0>();
    return amax;
}

template <bool VAR>
__global__ void __launch_bounds__(256, 2) k_residual_iso(const Lev L, const double* __restrict__ phi,
                                                         const double* __restrict__ rhs, double* __restrict__ res, int TZ,
                                                         double* __restrict__ norm_partial)
{
    extern __shared__ __align__(16) double smem[];
    const int i0 = blockIdx.x * SM_TX, j0 = blockIdx.y * SM_TY;
    const bool full = (i0 + SM_TX <= L.nn[0]) && (j0 + SM_TY <= L.nn[1]);
    double amax = full ? resid_iso_body<VAR, true>(L, phi, rhs, res, TZ, smem) : resid_iso_body<VAR, false>(L, phi, rhs, res, TZ, smem);
    if (norm_partial) {
        __syncthreads();
        double* sh = smem;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = amax;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int w = 0; w < 8; ++w) m = fmax(m, sh[w]);
            norm_partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = m;
        }
    }
}

}  // namespace b200np_dev
