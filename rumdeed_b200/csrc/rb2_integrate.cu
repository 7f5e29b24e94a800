// rb2_integrate.cu -- the O(N) kernels of the time step: Beeman position / velocity update
// with boundary + plane checks and Ramo current, particle add / mark / stable compaction,
// ordered event lists for the host writers, and the FP64 peak micro-benchmark.
//
// Replaces reference src/mod_verlet.F90:197-232 (Update_ElecIon_Position), :325-367
// (Check_Boundary_Planar, Check_Planes), :449-509 (Update_ElecIon_Velocity),
// src/mod_emission_tip.f90:1627-1647 (Check_Boundary_Tip) and src/mod_pair.F90:29-159,
// :169-339, :352-562 (Add_Particle, Mark_Particles_Remove, Remove_Particles).
//
// These kernels are HBM-bound.  The Beeman arithmetic uses explicitly rounded operations
// (__dmul_rn / __dadd_rn, never contracted into FMAs) in the reference's evaluation order,
// so positions -- and therefore the removal decisions z < 0, z > box_dim(3) -- are
// bit-identical to a non-contracted CPU evaluation of the same source expression.
#include "rb2_internal.cuh"

namespace {

constexpr int TPB = 256;
constexpr int SCAN_ITEMS = 1024;  // elements per scan block (256 threads x 4)

__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }

// ---- pack / unpack between the Fortran (3,N)+charge layout and pq ---------------------------
__global__ void k_pack(const double *__restrict__ pos3, const double *__restrict__ q, int n, double4 *__restrict__ pq)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pq[i] = make_double4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], q[i]);
}
__global__ void k_unpack(const double4 *__restrict__ pq, int n, double *__restrict__ pos3, double *__restrict__ q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = pq[i];
    if (pos3) { pos3[3 * i] = p.x; pos3[3 * i + 1] = p.y; pos3[3 * i + 2] = p.z; }
    if (q) q[i] = p.w;
}

// eta_coor, src/mod_hyperboloid_tip.f90:64-69 (explicitly rounded, reference order)
__device__ __forceinline__ double eta_coor(const TipParams &T, double x, double y, double z)
{
    const double zp = sub_(add_(z, T.a_foci), T.shift_z);
    const double zm = sub_(sub_(z, T.a_foci), T.shift_z);
    const double xy = add_(mul_(x, x), mul_(y, y));
    const double rp = __dsqrt_rn(add_(xy, mul_(zp, zp)));
    const double rm = __dsqrt_rn(add_(xy, mul_(zm, zm)));
    return mul_(__ddiv_rn(1.0, mul_(2.0, T.a_foci)), sub_(rp, rm));
}

// ---- Beeman position update + boundary + planes ---------------------------------------------
__global__ void __launch_bounds__(TPB)
k_update_position(int n, DevArrays A, int *__restrict__ mask, unsigned char *__restrict__ evcnt,
                  unsigned short *__restrict__ evbits, DevCounters *__restrict__ C, StepParams P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sp = A.species[i];
    if (sp == RB2_SPECIES_ATOM) {  // src/mod_verlet.F90:207
        evcnt[i] = 0;
        evbits[i] = 0;
        return;
    }
    double4 p = A.pq[i];
    double pos[3] = {p.x, p.y, p.z};
    double prev_z = p.z;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int e = 3 * i + c;
        const double v = A.vel[e], a = A.acc[e], ap = A.acc_prev[e];
        A.prev_pos[e] = pos[c];
        // pos + vel*dt + 1/6*(4*a - a_prev)*dt2, src/mod_verlet.F90:217-218
        const double t1 = add_(pos[c], mul_(v, P.dt));
        const double t2 = mul_(mul_(1.0 / 6.0, sub_(mul_(4.0, a), ap)), P.dt2);
        pos[c] = add_(t1, t2);
        A.acc_prev2[e] = ap;
        A.acc_prev[e] = a;
        A.acc[e] = 0.0;
    }
    // ptr_Check_Boundary
    int reason = 0;
    if (pos[2] < 0.0) reason = RB2_REMOVE_BOT;
    else if (pos[2] > P.box_z) reason = RB2_REMOVE_TOP;
    if (P.geometry == RB2_GEOM_TIP && reason == 0) {
        if (eta_coor(P.tip, pos[0], pos[1], pos[2]) < P.tip.eta_1) reason = RB2_REMOVE_BOT;
    }
    int bits = 0, cnt = 0;
    if (reason != 0 && mask[i] != 0) {  // Mark_Particles_Remove, src/mod_pair.F90:169-339
        mask[i] = 0;
        p.w = 0.0;
        atomicAdd(&C->mark_part, 1);
        if (sp == RB2_SPECIES_ELEC) {
            atomicAdd(&C->mark_elec, 1);
            if (reason == RB2_REMOVE_TOP) { atomicAdd(&C->top_part, 1); atomicAdd(&C->top_elec, 1); bits |= 1; }
            else { atomicAdd(&C->bot_part, 1); atomicAdd(&C->bot_elec, 1); bits |= 2; }
            cnt += 1;
        } else if (sp == RB2_SPECIES_ION) {
            atomicAdd(&C->mark_ion, 1);
            if (reason == RB2_REMOVE_TOP) { atomicAdd(&C->top_part, 1); atomicAdd(&C->top_ion, 1); }
            else { atomicAdd(&C->bot_part, 1); atomicAdd(&C->bot_ion, 1); }
        }
    }
    // Check_Planes, src/mod_verlet.F90:343-367
    for (int k = 0; k < P.planes_N; ++k) {
        const double zp = P.planes_z[k];
        if (zp > 0.0 && pos[2] > zp && prev_z < zp) { bits |= (4 << k); cnt += 1; }
    }
    A.pq[i] = make_double4(pos[0], pos[1], pos[2], p.w);
    evcnt[i] = (unsigned char)cnt;
    evbits[i] = (unsigned short)bits;
    if (cnt) atomicAdd(&C->n_events, cnt);
}

// ---- exclusive scan (block-local + block sums) ------------------------------------------------
struct ValEv { const unsigned char *c; __device__ int operator()(int i) const { return c[i]; } };
struct ValAlive { const int *m; __device__ int operator()(int i) const { return m[i] ? 1 : 0; } };

// `guard`: when given, the pass is skipped while *guard == 0 (the event passes of a step are always queued and
// usually have nothing to do)
template <class V>
__global__ void __launch_bounds__(TPB) k_scan_local(int n, V val, int *__restrict__ prefix, int *__restrict__ blocksum,
                                                    const int *__restrict__ guard)
{
    __shared__ int warp_tot[TPB / 32];
    if (guard && *guard == 0) return;
    const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
    int v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? val(base + k) : 0; s += v[k]; }
    // inclusive warp scan of the per-thread sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += warp_tot[k];
    int run = woff + inc - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < n) prefix[base + k] = run; run += v[k]; }
    if (threadIdx.x == TPB - 1) blocksum[blockIdx.x] = woff + inc;
}
// single block: exclusive scan of blocksum[0..nb) in place, total to *total
__global__ void __launch_bounds__(1024) k_scan_sums(int nb, int *__restrict__ blocksum, int *__restrict__ total,
                                                    const int *__restrict__ guard)
{
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (guard && *guard == 0) return;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int idx = b0 + threadIdx.x;
        const int s = (idx < nb) ? blocksum[idx] : 0;
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        int woff = 0;
        for (int k = 0; k < w; ++k) woff += warp_tot[k];
        const int carry = carry_s;
        if (idx < nb) blocksum[idx] = carry + woff + inc - s;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// ---- ordered event records ----------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_events_scatter(int n, DevArrays A, const unsigned char *__restrict__ evcnt, const unsigned short *__restrict__ evbits,
                 const int *__restrict__ prefix, const int *__restrict__ blocksum, rb2_event *__restrict__ out, int cap,
                 int planes_N, const double *__restrict__ vel, const int *__restrict__ guard)
{
    if (guard && *guard == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (evcnt[i] == 0) return;
    int o = prefix[i] + blocksum[i / SCAN_ITEMS];
    const int bits = evbits[i];
    const double4 p = A.pq[i];
    rb2_event e;
    e.index = i;
    e.x = p.x / rb2k::length_scale;
    e.y = p.y / rb2k::length_scale;
    e.vx = vel[3 * i]; e.vy = vel[3 * i + 1]; e.vz = vel[3 * i + 2];
    e.emit = A.emitter[i]; e.sec = A.section[i]; e.id = A.id[i];
    if (bits & 3) {
        e.kind = (bits & 1) ? 1 : 2;
        e.plane = -1;
        if (o < cap) out[o] = e;
        ++o;
    }
    for (int k = 0; k < planes_N; ++k) {
        if (bits & (4 << k)) {
            e.kind = 3;
            e.plane = k;
            if (o < cap) out[o] = e;
            ++o;
        }
    }
}

// ---- Beeman velocity update + Ramo current + velocity sums -----------------------------------
constexpr int NRED = 13;  // ramo[0..3], part(3), elec(3), ion(3)

// q * (v . E_zunit(pos)): the Shockley-Ramo term of one particle, src/mod_verlet.F90:481-488
__device__ __forceinline__ double ramo_term(const StepParams &P, const double4 &p, const double v[3])
{
    double ex = 0.0, ey = 0.0, ez;
    if (P.geometry == RB2_GEOM_TIP) {  // E_zunit_tip, src/mod_emission_tip.f90:133-141
        rb2_tip_field_E(P.tip, p.x, p.y, p.z, ex, ey, ez);
        ex = ex * P.tip.unit_scale_num / P.tip.unit_scale_den;
        ey = ey * P.tip.unit_scale_num / P.tip.unit_scale_den;
        ez = ez * P.tip.unit_scale_num / P.tip.unit_scale_den;
    } else {  // E_zunit_planar, src/mod_field_emission_v2.F90:158-166
        ez = -1.0 / P.d;
    }
    const double EzV = v[0] * ex + v[1] * ey + v[2] * ez;
    return p.w * EzV;
}

__global__ void __launch_bounds__(TPB)
k_update_velocity(int n, DevArrays A, StepParams P, double *__restrict__ redpart, const unsigned char *__restrict__ evcnt,
                  double *__restrict__ vel_save)
{
    __shared__ double sm[TPB / 32][NRED];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double r[NRED];
#pragma unroll
    for (int k = 0; k < NRED; ++k) r[k] = 0.0;
    if (i < n) {
        const int sp = A.species[i];
        if (sp != RB2_SPECIES_ATOM) {
            double v[3];
            // a particle with records keeps its pre-update velocity (what the records hold) in the spare velocity
            // array, so the record list can be rebuilt when it did not fit its buffer (rb2_rebuild_events)
            const bool keep = evcnt[i] != 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int e = 3 * i + c;
                if (keep) vel_save[e] = A.vel[e];
                // vel + 1/6*(2*a + 5*a_prev - a_prev2)*dt, src/mod_verlet.F90:470-473
                const double s = sub_(add_(mul_(2.0, A.acc[e]), mul_(5.0, A.acc_prev[e])), A.acc_prev2[e]);
                v[c] = add_(A.vel[e], mul_(mul_(1.0 / 6.0, s), P.dt));
                A.vel[e] = v[c];
            }
            const double4 p = A.pq[i];
            if (sp >= 0 && sp < 4) r[sp] = ramo_term(P, p, v);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[4 + c] = v[c];
                if (sp == RB2_SPECIES_ELEC) r[7 + c] = v[c];
                else if (sp == RB2_SPECIES_ION) r[10 + c] = v[c];
            }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
        double x = r[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[w][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NRED) {
        double x = 0.0;
        for (int k = 0; k < TPB / 32; ++k) x += sm[k][threadIdx.x];
        redpart[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = x;
    }
}
// one block per reduced quantity: fixed-order tree over the per-block partials
__global__ void __launch_bounds__(TPB) k_reduce_final(int nblocks, const double *__restrict__ redpart, double *__restrict__ out)
{
    __shared__ double sm[TPB];
    const double *src = redpart + (size_t)blockIdx.x * nblocks;
    double x = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += TPB) x += src[k];
    sm[threadIdx.x] = x;
    __syncthreads();
    for (int o = TPB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// ---- per-section Ramo current: ramo_current_emit(sec, emit), src/mod_verlet.F90:489-492 -------------------------
// A keyed sum with a FIXED order (bit-identical from run to run and on every replica): block b owns a contiguous
// range of particles and walks it tile by tile; inside a tile a warp groups its lanes by key (match.any), the lowest
// lane of each group adds the group's terms in ascending lane order, and the warps fold their group sums into the
// block's shared-memory table one warp after the other.  The per-block tables are added in block order by
// k_ramo_sections_final.  nkeys = n_sec * n_emit doubles of dynamic shared memory (<= 72 KB at MAX_SECTIONS = 9216).
__global__ void __launch_bounds__(TPB)
k_ramo_sections(int n, int per_block, DevArrays A, StepParams P, int n_sec, int n_emit, double *__restrict__ part)
{
    extern __shared__ double tab[];
    const int nkeys = n_sec * n_emit;
    for (int k = threadIdx.x; k < nkeys; k += TPB) tab[k] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i0 = blockIdx.x * per_block;
    const int i1 = min(n, i0 + per_block);
    for (int base = i0; base < i1; base += TPB) {  // block-uniform trip count
        const int i = base + threadIdx.x;
        int key = -1;
        double x = 0.0;
        if (i < i1) {
            const int sp = A.species[i];
            const int sec = A.section[i], emit = A.emitter[i];
            if (sp != RB2_SPECIES_ATOM && sec >= 1 && sec <= n_sec && emit >= 1 && emit <= n_emit) {
                key = (emit - 1) * n_sec + (sec - 1);
                const double v[3] = {A.vel[3 * i], A.vel[3 * i + 1], A.vel[3 * i + 2]};
                x = ramo_term(P, A.pq[i], v);
            }
        }
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        const bool leader = (key >= 0) && (lane == __ffs(grp) - 1);
        double sum = 0.0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const double t = __shfl_sync(0xffffffffu, x, l);
            if ((grp >> l) & 1u) sum += t;
        }
        for (int ww = 0; ww < TPB / 32; ++ww) {
            if (w == ww && leader) tab[key] += sum;
            __syncthreads();
        }
    }
    for (int k = threadIdx.x; k < nkeys; k += TPB) part[(size_t)blockIdx.x * nkeys + k] = tab[k];
}
__global__ void __launch_bounds__(TPB) k_ramo_sections_final(int nblocks, int nkeys, const double *__restrict__ part, double *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    double x = 0.0;
    for (int b = 0; b < nblocks; ++b) x += part[(size_t)b * nkeys + k];
    out[k] = x;
}

// ---- stable compaction ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_compact(int n, int step, DevArrays S, DevArrays D, const int *__restrict__ mask, const int *__restrict__ prefix,
          const int *__restrict__ blocksum, unsigned long long *__restrict__ life_hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!mask[i]) {  // record_lifetime, src/mod_pair.F90:1135-1160
        int lt = step - S.step[i];
        if (lt <= 0) lt = 1;
        if (lt > RB2_MAX_LIFE_TIME) lt = RB2_MAX_LIFE_TIME;
        const int s = S.species[i];
        if (s >= 1 && s <= 3) atomicAdd(&life_hist[lt * 4 + s], 1ull);
        return;
    }
    const int j = prefix[i] + blocksum[i / SCAN_ITEMS];
    D.pq[j] = S.pq[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        D.prev_pos[3 * j + c] = S.prev_pos[3 * i + c];
        D.vel[3 * j + c] = S.vel[3 * i + c];
        D.acc[3 * j + c] = S.acc[3 * i + c];
        D.acc_prev[3 * j + c] = S.acc_prev[3 * i + c];
        D.acc_prev2[3 * j + c] = S.acc_prev2[3 * i + c];
    }
    D.mass[j] = S.mass[i];
    D.species[j] = S.species[i];
    D.step[j] = S.step[i];
    D.emitter[j] = S.emitter[i];
    D.section[j] = S.section[i];
    D.life[j] = S.life[i];
    D.id[j] = S.id[i];
}

__global__ void k_fill_int(int *p, int n, int v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_iota(int *p, int n, int v0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v0 + i;
}

// ---- Add_Particle, src/mod_pair.F90:29-159 ------------------------------------------------------
__global__ void k_add(int k, int slot0, int id0, int step, const double *__restrict__ pos, const double *__restrict__ vel,
                      const int *__restrict__ species, const int *__restrict__ emit, const int *__restrict__ sec,
                      const int *__restrict__ life, DevArrays A, int *__restrict__ mask, StepParams P)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const int s = slot0 + t;
    const int sp = species[t];
    double q, m;
    if (sp == RB2_SPECIES_ELEC) { q = -1.0 * rb2k::q_0; m = 1.0 * rb2k::m_0; }
    else if (sp == RB2_SPECIES_ION) { q = +1.0 * rb2k::q_0; m = rb2k::m_N2p; }
    else { q = 0.0; m = rb2k::m_N2; }
    const double x = pos[3 * t], y = pos[3 * t + 1], z = pos[3 * t + 2];
    A.pq[s] = make_double4(x, y, z, q);
    A.mass[s] = m;
    // seed the Beeman history with the vacuum-field acceleration, :133-139
    double fx = 0.0, fy = 0.0, fz = P.pl.E_z;
    if (P.geometry == RB2_GEOM_TIP) rb2_tip_field_E(P.tip, x, y, z, fx, fy, fz);
    const double f = q / m;
    const double a3[3] = {f * fx, f * fy, f * fz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A.prev_pos[3 * s + c] = -1.0 * rb2k::length_scale;
        A.vel[3 * s + c] = vel[3 * t + c];
        A.acc[3 * s + c] = a3[c];
        A.acc_prev[3 * s + c] = a3[c];
        A.acc_prev2[3 * s + c] = a3[c];
    }
    A.species[s] = sp;
    A.step[s] = step;
    A.emitter[s] = emit[t];
    A.section[s] = sec[t];
    A.life[s] = life[t];
    A.id[s] = id0 + t;
    mask[s] = 1;
}

// ---- Mark_Particles_Remove for host-chosen particles ------------------------------------------
__global__ void k_mark(int k, const int *__restrict__ index, const int *__restrict__ reason, int n, DevArrays A,
                       int *__restrict__ mask, DevCounters *__restrict__ C)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const int i = index[t];
    if (i < 0 || i >= n) return;
    const int sp = A.species[i];
    if (sp != RB2_SPECIES_ELEC && sp != RB2_SPECIES_ION && sp != RB2_SPECIES_ATOM) return;
    if (atomicExch(&mask[i], 0) == 0) return;  // already marked: no-op
    A.pq[i].w = 0.0;
    atomicAdd(&C->mark_part, 1);
    const int r = reason[t];
    if (sp == RB2_SPECIES_ELEC) {
        atomicAdd(&C->mark_elec, 1);
        if (r == RB2_REMOVE_TOP) { atomicAdd(&C->top_part, 1); atomicAdd(&C->top_elec, 1); }
        else if (r == RB2_REMOVE_BOT) { atomicAdd(&C->bot_part, 1); atomicAdd(&C->bot_elec, 1); }
        else if (r == RB2_REMOVE_RECOM) { atomicAdd(&C->recom_part, 1); atomicAdd(&C->recom_elec, 1); }
        else if (r == RB2_REMOVE_ION) { atomicAdd(&C->ion_part, 1); atomicAdd(&C->ion_elec, 1); }
    } else if (sp == RB2_SPECIES_ION) {
        atomicAdd(&C->mark_ion, 1);
        if (r == RB2_REMOVE_TOP) { atomicAdd(&C->top_part, 1); atomicAdd(&C->top_ion, 1); }
        else if (r == RB2_REMOVE_BOT) { atomicAdd(&C->bot_part, 1); atomicAdd(&C->bot_ion, 1); }
        else if (r == RB2_REMOVE_RECOM) { atomicAdd(&C->recom_part, 1); atomicAdd(&C->recom_ion, 1); }
    } else {
        atomicAdd(&C->mark_atom, 1);
        if (r == RB2_REMOVE_ION) { atomicAdd(&C->ion_part, 1); atomicAdd(&C->ion_atom, 1); }
    }
}

// ---- FP64 peak: 8 independent DFMA chains per thread -----------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double *__restrict__ sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1.0e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

inline int nblk(int n) { return (n + TPB - 1) / TPB; }

template <class V>
int run_scan(Rb2Ctx &ctx, int n, V val, const int *guard = nullptr)
{
    const int nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    k_scan_local<V><<<nb, TPB, 0, ctx.stream>>>(n, val, ctx.prefix, ctx.blocksum, guard);
    RB2_CUDA(cudaGetLastError());
    k_scan_sums<<<1, 1024, 0, ctx.stream>>>(nb, ctx.blocksum, ctx.d_total, guard);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(2);
    return RB2_OK;
}

}  // namespace

int rb2_launch_pack(Rb2Ctx &ctx, const double *pos3, const double *q, int n, double4 *pq)
{
    if (n < 1) return RB2_OK;
    k_pack<<<nblk(n), TPB, 0, ctx.stream>>>(pos3, q, n, pq);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}
int rb2_launch_unpack(Rb2Ctx &ctx, const double4 *pq, int n, double *pos3, double *q)
{
    if (n < 1) return RB2_OK;
    k_unpack<<<nblk(n), TPB, 0, ctx.stream>>>(pq, n, pos3, q);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

int rb2_launch_update_position(Rb2Ctx &ctx)
{
    if (ctx.n < 1) return RB2_OK;
    const StepParams P = rb2_make_step_params(ctx.cfg);
    k_update_position<<<nblk(ctx.n), TPB, 0, ctx.stream>>>(ctx.n, ctx.a, ctx.mask, ctx.evcnt, ctx.evbits, ctx.d_counters, P);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

namespace {
int ensure_events(Rb2Ctx &ctx, int want)
{
    if (want > ctx.ev_cap) {
        if (ctx.d_events) RB2_CUDA(cudaFree(ctx.d_events));
        ctx.d_events = nullptr;
        ctx.ev_cap = 0;
        RB2_CUDA(cudaMalloc(&ctx.d_events, (size_t)want * sizeof(rb2_event)));
        ctx.ev_cap = want;
    }
    return RB2_OK;
}

}  // namespace

// Queue the ordered record list of the position update that is in the stream (absorbed electrons, plane crossings,
// ascending particle index) into ctx.d_events.  The passes read the record count on the device and do nothing when
// it is zero, so the step never waits for the host; records beyond the buffer are dropped here and rebuilt by
// rb2_rebuild_events once the host knows the count.
int rb2_launch_events(Rb2Ctx &ctx)
{
    if (ctx.n < 1) return RB2_OK;
    int rc = ensure_events(ctx, ctx.ev_min);
    if (rc != RB2_OK) return rc;
    const int *guard = &ctx.d_counters->n_events;
    rc = run_scan(ctx, ctx.n, ValEv{ctx.evcnt}, guard);
    if (rc != RB2_OK) return rc;
    k_events_scatter<<<nblk(ctx.n), TPB, 0, ctx.stream>>>(ctx.n, ctx.a, ctx.evcnt, ctx.evbits, ctx.prefix, ctx.blocksum,
                                                         ctx.d_events, ctx.ev_cap, ctx.cfg.planes_N, ctx.a.vel, guard);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

// The record list did not fit: grow the buffer and scatter again (the scan of the step is still in place).
// vel: the velocities the records must hold -- the current ones before the velocity update, the saved ones after.
int rb2_rebuild_events(Rb2Ctx &ctx, int n_events, bool after_velocity_update)
{
    int rc = ensure_events(ctx, n_events + n_events / 2 + 1024);
    if (rc != RB2_OK) return rc;
    k_events_scatter<<<nblk(ctx.n), TPB, 0, ctx.stream>>>(ctx.n, ctx.a, ctx.evcnt, ctx.evbits, ctx.prefix, ctx.blocksum,
                                                         ctx.d_events, ctx.ev_cap, ctx.cfg.planes_N,
                                                         after_velocity_update ? ctx.b.vel : ctx.a.vel, nullptr);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

int rb2_launch_update_velocity(Rb2Ctx &ctx)
{
    const int nb = ctx.n > 0 ? nblk(ctx.n) : 0;
    if (nb > ctx.redpart_blocks) {
        if (ctx.d_redpart) RB2_CUDA(cudaFree(ctx.d_redpart));
        ctx.d_redpart = nullptr;
        ctx.redpart_blocks = 0;
        const int want = nb + nb / 2 + 64;
        RB2_CUDA(cudaMalloc(&ctx.d_redpart, (size_t)want * NRED * sizeof(double)));
        ctx.redpart_blocks = want;
    }
    if (nb == 0) {
        RB2_CUDA(cudaMemsetAsync(ctx.d_red, 0, 16 * sizeof(double), ctx.stream));
        return RB2_OK;
    }
    const StepParams P = rb2_make_step_params(ctx.cfg);
    k_update_velocity<<<nb, TPB, 0, ctx.stream>>>(ctx.n, ctx.a, P, ctx.d_redpart, ctx.evcnt, ctx.b.vel);
    RB2_CUDA(cudaGetLastError());
    k_reduce_final<<<NRED, TPB, 0, ctx.stream>>>(nb, ctx.d_redpart, ctx.d_red);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(2);
    return RB2_OK;
}

// ramo_current_emit of the velocities now in the store (queued behind the velocity update); the result is copied to
// ctx.h_ramo_sec in stream order.
int rb2_launch_ramo_sections(Rb2Ctx &ctx)
{
    const int nkeys = ctx.ramo_n_sec * ctx.ramo_n_emit;
    if (nkeys < 1) return RB2_OK;
    if (!ctx.d_ramo_sec) {
        RB2_CUDA(cudaMalloc(&ctx.d_ramo_sec, (size_t)nkeys * sizeof(double)));
        RB2_CUDA(cudaMallocHost(&ctx.h_ramo_sec, (size_t)nkeys * sizeof(double)));
        ctx.ramo_blocks = 2 * ctx.sm_count;
        RB2_CUDA(cudaMalloc(&ctx.d_ramo_part, (size_t)ctx.ramo_blocks * nkeys * sizeof(double)));
        RB2_CUDA(cudaFuncSetAttribute(k_ramo_sections, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(nkeys * sizeof(double))));
    }
    if (ctx.n < 1) {
        RB2_CUDA(cudaMemsetAsync(ctx.d_ramo_sec, 0, (size_t)nkeys * sizeof(double), ctx.stream));
    } else {
        int nb = nblk(ctx.n);
        if (nb > ctx.ramo_blocks) nb = ctx.ramo_blocks;
        int per_block = (ctx.n + nb - 1) / nb;
        per_block = (per_block + TPB - 1) / TPB * TPB;
        nb = (ctx.n + per_block - 1) / per_block;
        const StepParams P = rb2_make_step_params(ctx.cfg);
        k_ramo_sections<<<nb, TPB, (size_t)nkeys * sizeof(double), ctx.stream>>>(ctx.n, per_block, ctx.a, P, ctx.ramo_n_sec, ctx.ramo_n_emit, ctx.d_ramo_part);
        RB2_CUDA(cudaGetLastError());
        k_ramo_sections_final<<<(nkeys + TPB - 1) / TPB, TPB, 0, ctx.stream>>>(nb, nkeys, ctx.d_ramo_part, ctx.d_ramo_sec);
        RB2_CUDA(cudaGetLastError());
        RB2_LAUNCHED(2);
    }
    RB2_CUDA(cudaMemcpyAsync(ctx.h_ramo_sec, ctx.d_ramo_sec, (size_t)nkeys * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    return RB2_OK;
}

// Gather the survivors (stable) into the spare array set and swap; bins the life times.
int rb2_launch_compact(Rb2Ctx &ctx, int step)
{
    if (ctx.n < 1) return RB2_OK;
    int rc = run_scan(ctx, ctx.n, ValAlive{ctx.mask});
    if (rc != RB2_OK) return rc;
    k_compact<<<nblk(ctx.n), TPB, 0, ctx.stream>>>(ctx.n, step, ctx.a, ctx.b, ctx.mask, ctx.prefix, ctx.blocksum, ctx.life_hist);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    DevArrays t = ctx.a;
    ctx.a = ctx.b;
    ctx.b = t;
    return RB2_OK;
}

int rb2_launch_add(Rb2Ctx &ctx, int k, int slot0, int id0, int step, const double *d_pos, const double *d_vel,
                   const int *d_species, const int *d_emit, const int *d_sec, const int *d_life)
{
    if (k < 1) return RB2_OK;
    const StepParams P = rb2_make_step_params(ctx.cfg);
    k_add<<<nblk(k), TPB, 0, ctx.stream>>>(k, slot0, id0, step, d_pos, d_vel, d_species, d_emit, d_sec, d_life, ctx.a, ctx.mask, P);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

int rb2_launch_mark(Rb2Ctx &ctx, int k, const int *d_index, const int *d_reason)
{
    if (k < 1) return RB2_OK;
    k_mark<<<nblk(k), TPB, 0, ctx.stream>>>(k, d_index, d_reason, ctx.n, ctx.a, ctx.mask, ctx.d_counters);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

int rb2_launch_fill_defaults(Rb2Ctx &ctx, int n, bool species, bool step, bool emitter, bool section, bool life, bool id)
{
    if (n < 1) return RB2_OK;
    cudaStream_t st = ctx.stream;
    int cnt = 0;
    if (species) { k_fill_int<<<nblk(n), TPB, 0, st>>>(ctx.a.species, n, RB2_SPECIES_ELEC); ++cnt; }
    if (step) { k_fill_int<<<nblk(n), TPB, 0, st>>>(ctx.a.step, n, 0); ++cnt; }
    if (emitter) { k_fill_int<<<nblk(n), TPB, 0, st>>>(ctx.a.emitter, n, 1); ++cnt; }
    if (section) { k_fill_int<<<nblk(n), TPB, 0, st>>>(ctx.a.section, n, 1); ++cnt; }
    if (life) { k_fill_int<<<nblk(n), TPB, 0, st>>>(ctx.a.life, n, -1); ++cnt; }
    if (id) { k_fill_iota<<<nblk(n), TPB, 0, st>>>(ctx.a.id, n, 0); ++cnt; }
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(cnt);
    return RB2_OK;
}

int rb2_launch_fill_mask(Rb2Ctx &ctx, int n)
{
    if (n < 1) return RB2_OK;
    k_fill_int<<<nblk(n), TPB, 0, ctx.stream>>>(ctx.mask, n, 1);
    RB2_CUDA(cudaGetLastError());
    RB2_LAUNCHED(1);
    return RB2_OK;
}

int rb2_launch_fp64_peak(Rb2Ctx &ctx, double ms_target, double *tflops, float *ms_out)
{
    double *sink = nullptr;
    RB2_CUDA(cudaMalloc(&sink, sizeof(double)));
    const int blocks = ctx.sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    RB2_CUDA(cudaEventCreate(&e0));
    RB2_CUDA(cudaEventCreate(&e1));
    int iters = 2000;
    float ms = 0.f;
    for (int attempt = 0; attempt < 6; ++attempt) {
        k_fp64_peak<<<blocks, threads, 0, ctx.stream>>>(iters / 4 + 1, 1.0, sink);  // warm-up
        RB2_CUDA(cudaEventRecord(e0, ctx.stream));
        k_fp64_peak<<<blocks, threads, 0, ctx.stream>>>(iters, 1.0, sink);
        RB2_CUDA(cudaEventRecord(e1, ctx.stream));
        RB2_CUDA(cudaEventSynchronize(e1));
        RB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        RB2_LAUNCHED(2);
        if (ms >= 0.6 * ms_target || iters > (1 << 24)) break;
        const double scale = (ms > 1e-3) ? (ms_target / ms) : 16.0;
        iters = (int)(iters * (scale > 16.0 ? 16.0 : scale)) + 1;
    }
    const double flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
    *tflops = flops / (ms * 1.0e-3) / 1.0e12;
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return RB2_OK;
}
