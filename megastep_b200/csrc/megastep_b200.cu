// megastep_b200.cu — sm_100a kernels + C ABI (include/megastep_b200.h) for megastep's per-step hot path.
//
// Replaces, from scratch, the reference's megastep/src/kernels.cu:
//   physics()  = collision_kernel (:179-210) + ~15 ATen elementwise launches (:223-227)        -> physics_kernel
//   render()   = draw_kernel (:297-318) + raycast_kernel (:326-383) + shader_kernel (:407-450)  -> view_kernel + dyn_kernel
//   msb_step() = MomentumMovement (modules.py:106-118) + physics + render + RGB/Depth/IMU heads -> the same three launches
//   bake()     = baking_kernel (:270-284)                                                       -> bake_kernel
//
// One-off, per scenery: table_kernel (msb_build_table), vis_kernel (msb_build_visibility), bake_kernel (msb_bake).
//
// Layout / mapping:
//   * every scenery carries a SPATIAL TABLE built once: each env's static segments packed sort-tile-recursive into runs
//     of 16 with one bounding box per run, padded per env, plus per row the texel offset / count and line index;
//   * view_kernel: one CTA per env stages the env's table with three 1-D bulk (TMA) copies on one mbarrier; one warp
//     per (agent, block of rays) visits the run boxes nearest first, bins 32 segments at a time (lane = segment),
//     tests them lane = ray on the candidates a ballot yields, and never opens boxes hidden behind what it already
//     hit; the reference's order-dependent nearest-hit rule is restored exactly by replaying near-tied rays;
//   * the ray/line cosine and its sqrt are computed for the winning line only (the reference does it for every line);
//   * rays that hit another agent need the dynamic light at the hit point (I lights x W occluders): queued as windows
//     of 4 pixels for dyn_kernel, which spreads them over the whole GPU, consumes the queue while view_kernel still
//     fills it (programmatic dependent launch) and scans only the runs near each light ray that neither the
//     light-visibility grid nor the remembered occluders settle;
//   * physics_kernel: one warp per agent, box-culled over the same table, fused movement / integration; it also
//     prefetches each env's table into L2 for the view_kernel behind it.
//
// No tensor cores: nothing on this path is a dense contraction.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>

#include "../../include/megastep_b200.h"
#include "msb_math.cuh"

using namespace msb;

// ---------------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------------
struct KArgs {
    msb_params p;
    msb_scenery s;
    msb_agents a;
    float* progress;
    msb_render_out out;
    msb_obs_out obs;
    msb_movement mv;
    int32_t has_obs;
    int32_t stage_rec;      // view_kernel stages the rows' records too (worth their shared memory when there are many rays per env)
    int32_t idx32;          // every output index (3 * N * A * R at most) fits 32 bits
    int32_t out_mask;       // which outputs are wanted (OUT_* bits): one uniform test instead of a 64-bit pointer compare each
    int32_t sub_shift;      // log2(obs.subsample)
    int32_t has_mv;
    int32_t ray_blocks;     // RB: warps per agent in view_kernel
    int32_t rb_shift;       // log2(RB) when RB is a power of two, else -1
    int32_t seg_cap;        // bake_kernel: float4 slots for an env's lines in shared memory (>= max_lines)
    int32_t wcap;           // view_kernel: slots for the env's padded run table (multiple of 16)
    float inv_fps;          // IEEE 1/fps (ATen's tensor/scalar == tensor*(1/scalar), kernels.cu:224,226)
    float mv_keep, mv_dv, mv_dw;   // 1-decay, accel/fps, ang_accel/fps evaluated in double like the Python does
    float inv_max_depth, inv_speed, inv_ang, inv_sub;   // reciprocals ATen would multiply by
    float xclip;            // camera-space depth well inside every ray's near plane (culling only)
    unsigned long long* stats;  // optional diagnostics counters (may be null)
    int32_t debug_skip_dyn;     // timing experiments only: leave agent-hit rays unlit
    int32_t debug_no_vis;       // tests: ignore the visibility grid (same results, more shadow scans)
    const int32_t* env_order;   // view_kernel: env of CTA b (null: b)
    int32_t prefetch_view;      // physics_kernel: prefetch each env's table into L2 for the view_kernel that follows
    // queue of ray chunks whose agent-hit pixels are lit by dyn_kernel (load-balanced second pass)
    int* dyn_ctrl;              // [0] entries reserved, [1] CTAs of dyn_kernel done, [2] entries handed out beyond each CTA's first,
                                // [3] CTAs of view_kernel done
    unsigned char* dyn_entries; // null -> dynamic lights are resolved inline by the ray's own warp
    int32_t dyn_cap;            // entries that fit
    int32_t dyn_window;         // PS: pixels per entry = max(DYN_MIN_WINDOW, subsample); entry = DYN_HDR + 32 * PS bytes
    int* dyn_cache;             // [N][A][32] last occluder of each light as seen from (around) each agent; a hint
    int32_t view_ctas;          // what dyn_ctrl[3] counts up to: envs (= CTAs of view_kernel) that will publish nothing more
    // tick_kernel (the persistent form of view_kernel)
    int* sched;                 // [0] envs handed out beyond every CTA's first p_stages, [1] CTAs out, [2] a bounded wait ran out (bug); null: envs are dealt round-robin
    int32_t p_stages;           // S: envs staged per CTA at a time (ring of shared-memory stages)
    int32_t p_stage_bytes;      // bytes of one stage
    int32_t p_items;            // (agent, ray block) items per env = n_agents * ray_blocks
    int32_t dyn_groups;         // merged second pass: warps (tickets) sharing one queue entry, each owning the lights i = g (mod groups)
};

#ifndef MSB_DYN_BLOCKS
#define MSB_DYN_BLOCKS 7        // 72 registers
#endif
#ifndef MSB_VIEW_THREADS
#define MSB_VIEW_THREADS 256   // 256 threads x 4 blocks -> at most 64 registers per thread (launched with 128: 8 CTAs/SM)
#define MSB_VIEW_BLOCKS 4
#endif
enum { ST_ANG = 0, ST_PX = 1, ST_PY = 2, ST_AV = 3, ST_VX = 4, ST_VY = 5, ST_SN = 6, ST_CS = 7, ST_STRIDE = 8 };   // SN, CS: view_kernel only
enum { STAT_TESTS = 0, STAT_GROUPS = 1, STAT_DYN_RAYS = 2, STAT_DYN_ITERS = 3, STAT_DYN_ENTRIES = 4, STAT_REPLAYS = 5,
       STAT_DYN_CYCLES = 6, STAT_DYN_MAXCYC = 7, STAT_DYN_WARPMAX = 8, STAT_DYN_SLOW = 9, STAT_DYN_KERNEL = 10, STAT_DYN_SCANS = 11, STAT_DYN_SCANS_LIT = 12, STAT_DYN_ITERS_LIT = 13,
       // tick_kernel, warp-cycles summed over the grid: in ray items, waiting for a stage, in second-pass tickets, idle in the
       // drain, staging envs, whole kernel; counts of tickets run and items; the longest warp; waits that actually spun
       STAT_T_ITEMS = 16, STAT_T_WAIT = 17, STAT_T_DYN = 18, STAT_T_DRAIN = 19, STAT_T_PREP = 20, STAT_T_TOTAL = 21, STAT_N_DYN = 22,
       STAT_N_ITEMS = 23, STAT_T_MAXWARP = 24, STAT_N_SPUN = 25, STAT_T_FETCH = 26, STAT_SLOTS = 32 };
enum { OUT_INDICES = 1, OUT_LOCATIONS = 2, OUT_DOTS = 4, OUT_DISTANCES = 8, OUT_SCREEN = 16, OUT_RGB = 32, OUT_DEPTH = 64, OUT_IMU = 128 };
enum { VRUN = 16 };                       // segments per run of the spatial table
enum { DYN_MIN_WINDOW = 4 };               // pixels per queue entry: max(4, subsample) adjacent pixels (a 'window')
constexpr float VIS_CELL = 0.25f, VIS_INV_CELL = 4.f;   // the light-visibility grid's cell, metres
enum { DYN_HDR = 48, DYN_FLAG = 44 };      // bytes of a queue entry's header (32 bytes per pixel follow); offset of its 'published' word

// ---------------------------------------------------------------------------------------------------------------
// physics
// ---------------------------------------------------------------------------------------------------------------

// collision(p0, v0, p1, v1) — kernels.cu:119-133 (+ project :92-106), op order per docs/REFERENCE_ARITHMETIC.md
__device__ __forceinline__ float collide_agents(float p0x, float p0y, float m0x, float m0y, float p1x, float p1y,
                                                float m1x, float m1y, float rF, float r2) {
    const float Ux = ffma(m0x, rF, -fmul(m1x, rF));
    const float Uy = ffma(m0y, rF, -fmul(m1y, rF));
    const float ulen = sqrt_(ffma(Ux, Ux, fmul(Uy, Uy)));
    const float u = fadd(ulen, 1e-6f);
    const float PQx = fsub(p1x, p0x), PQy = fsub(p1y, p0y);
    const float s = fmul(dot2(Ux, PQx, Uy, PQy), rcp(fmul(u, u)));
    const float d = fmul(fabsf(cross2(Uy, PQx, Ux, PQy)), rcp(u));
    float x = 1.f;
    if ((s > 0.f) && (d < r2)) {
        const float back = sqrt_(ffma(-d, d, fmul(r2, r2)));
        x = fminf(x, sens(ffma(-back, rcp(ulen), s)));
    }
    return x;
}

// collision(p, v, l) — kernels.cu:135-171. v is already velocity/fps; vlen = |v|.
__device__ __forceinline__ float collide_line(float px, float py, float vx, float vy, float vlen, float u, float uu,
                                              float4 l, float r1, float r1sq) {
    const float Vx = fsub(l.z, l.x), Vy = fsub(l.w, l.y);
    float x = 1.f;

    // passing through l (:143-146)
    {
        const float PQx = fsub(l.x, px), PQy = fsub(l.y, py);
        const Hit mid = intersect_pre(vx, vy, Vx, Vy, PQx, PQy, cross2(Vy, PQx, Vx, PQy));
        if ((0.f < mid.s) && (mid.s < 1.f) && (0.f < mid.t) && (mid.t < 1.f)) {
            const float cr = cross2(Vy, fsub(px, l.x), Vx, fsub(py, l.y));
            const float uV = fadd(sqrt_(ffma(Vx, Vx, fmul(Vy, Vy))), 1e-6f);
            const float d = fmul(fabsf(cr), rcp(uV));
            x = fminf(x, sens(fmul(ffma(rcp(d), -r1, 1.f), mid.s)));
        }
    }
    // passing within r of l.a, then l.b (:149-160)
    const float ruu = rcp(uu), ru = rcp(u);
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const float ex = e ? l.z : l.x, ey = e ? l.w : l.y;
        const float PQx = fsub(ex, px), PQy = fsub(ey, py);
        const float s = fmul(dot2(vx, PQx, vy, PQy), ruu);
        const float d = fmul(fabsf(cross2(vy, PQx, vx, PQy)), ru);
        if ((0.f < s) && (d < r1)) {
            const float back = sqrt_(ffma(-d, d, r1sq));
            x = fminf(x, sens(ffma(-back, rcp(vlen), s)));
        }
    }
    // end point within r of the interior of l (:163-168)
    {
        const float PQx = fsub(fadd(px, vx), l.x), PQy = fsub(fadd(py, vy), l.y);
        const float uV = fadd(sqrt_(ffma(Vx, Vx, fmul(Vy, Vy))), 1e-6f);
        const float s = fmul(dot2(Vx, PQx, Vy, PQy), rcp(fmul(uV, uV)));
        const float rV = rcp(uV);
        const float dq = fmul(fabsf(cross2(Vy, PQx, Vx, PQy)), rV);
        if ((0.f < s) && (s < 1.f) && (dq < r1)) {
            const float cr = fabsf(cross2(Vy, fsub(px, l.x), Vx, fsub(py, l.y)));
            x = fminf(x, sens(fmul(ffma(cr, rV, -r1), rcp(ffma(cr, rV, -dq)))));
        }
    }
    return x;
}

// collision_kernel (kernels.cu:179-210) for one agent, by one warp, over the env's run table: lane b tests run b's box
// against the square the agent can reach this tick; only the overlapping runs are read, two per iteration (a run per
// half warp), and a segment runs the reference's circle-vs-segment test only if its own bounding box overlaps too.
// The minimum over obstacles is order-free. See DESIGN.md ("physics cull") for why the skipped ones cannot matter.
// SHARED: the table is staged in shared memory (view_kernel); otherwise it is read through the read-only path.
template <bool SHARED>
__device__ __forceinline__ float physics_agent(const float* st_in, int A, int a, int lane, const float4* occ,
                                               const float4* boxes, int W, int nb, float rF, float r1, float r2) {
    const float* me = st_in + a * ST_STRIDE;
    const float px = me[ST_PX], py = me[ST_PY], mx = me[ST_VX], my = me[ST_VY];
    const float vx = fmul(mx, rF), vy = fmul(my, rF);
    const float vlen = sqrt_(ffma(vx, vx, fmul(vy, vy)));
    const float r1sq = fmul(r1, r1);
    // slow but moving agents: project()'s +1e-6 distorts distances -> test everything; exactly stationary ones can
    // only trigger the end-point branch (:163-168: every other branch needs s > 0), which the same radius covers
    const bool can_cull = vlen >= 1e-3f || (vx == 0.f && vy == 0.f);
    const float rho = 1.05f * vlen + 2.2f * r1 + 0.02f;
    float x = 1.f;
    // other agents (:193-200): start-of-step state, no sequential resolution
    for (int d1 = lane; d1 < A; d1 += 32) {
        if (d1 != a) {
            const float* o = st_in + d1 * ST_STRIDE;
            x = fminf(x, collide_agents(px, py, mx, my, o[ST_PX], o[ST_PY], o[ST_VX], o[ST_VY], rF, r2));
        }
    }
    const float u = fadd(vlen, 1e-6f), uu = fmul(u, u);
    const int slot = lane / VRUN, within = lane - slot * VRUN;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        bool visit = false;
        if (b0 + lane < nb) {
            const float4 bx = SHARED ? boxes[b0 + lane] : __ldg(boxes + b0 + lane);
            visit = !(can_cull && (bx.x > px + rho || bx.z < px - rho || bx.y > py + rho || bx.w < py - rho));
        }
        unsigned runs = __ballot_sync(0xffffffffu, visit);
        while (runs) {
            const unsigned rest = runs & (runs - 1);
            const int n0 = __ffs(runs) - 1, n1 = rest ? __ffs(rest) - 1 : -1;
            const int nth = slot == 0 ? n0 : n1;
            const int l = nth >= 0 ? VRUN * (b0 + nth) + within : W;
            if (l < W) {
                const float4 s4 = SHARED ? occ[l] : __ldg(occ + l);
                const bool outside = (fminf(s4.x, s4.z) > px + rho) || (fmaxf(s4.x, s4.z) < px - rho) ||
                                     (fminf(s4.y, s4.w) > py + rho) || (fmaxf(s4.y, s4.w) < py - rho);
                if (!(can_cull && outside)) x = fminf(x, collide_line(px, py, vx, vy, vlen, u, uu, s4, r1, r1sq));
            }
            runs = rest & (rest - 1);
        }
    }
    return warp_min(x);
}

// The ATen epilogue of physics() (kernels.cu:223-227) for one agent: integrate, wrap the angle, kill the momentum of
// agents that hit something. Plain in-place stores (no storage swap). Leaves the new state in st_out.
__device__ __forceinline__ void physics_integrate(const KArgs& k, const float* st_in, float* st_out, int n, int a, float x,
                                                  bool want_sincos = false) {
    const float* me = st_in + a * ST_STRIDE;
    float* o = st_out + a * ST_STRIDE;
    const int64_t i = (int64_t)n * k.s.n_agents + a;
    const float npx = __fadd_rn(me[ST_PX], __fmul_rn(__fmul_rn(x, me[ST_VX]), k.inv_fps));
    const float npy = __fadd_rn(me[ST_PY], __fmul_rn(__fmul_rn(x, me[ST_VY]), k.inv_fps));
    float ang = __fadd_rn(me[ST_ANG], __fmul_rn(__fmul_rn(x, me[ST_AV]), k.inv_fps));
    ang = __fsub_rn(remainder_(__fadd_rn(remainder_(ang, 360.f), 180.f), 360.f), 180.f);
    const bool hit = x < 1.f;
    const float nvx = hit ? 0.f : me[ST_VX], nvy = hit ? 0.f : me[ST_VY], nav = hit ? 0.f : me[ST_AV];
    k.a.angles[i] = ang;
    reinterpret_cast<float2*>(k.a.positions)[i] = make_float2(npx, npy);
    k.a.angvelocity[i] = nav;
    reinterpret_cast<float2*>(k.a.velocity)[i] = make_float2(nvx, nvy);
    if (k.progress) k.progress[i] = x;
    o[ST_ANG] = ang; o[ST_PX] = npx; o[ST_PY] = npy; o[ST_AV] = nav; o[ST_VX] = nvx; o[ST_VY] = nvy;
    if (want_sincos) sincos_deg(ang, o[ST_SN], o[ST_CS]);
}

// MomentumMovement (modules.py:106-118) for one agent: velocities decay and take the chosen action's impulse.
__device__ __forceinline__ void momentum_movement(const KArgs& k, int act, float ang, float& av, float2& vel) {
    const float keep = k.mv_keep, dv = k.mv_dv, dw = k.mv_dw;
    // action table of modules.py:95-96: 0 noop, 1 +y, 2 -y, 3 +x, 4 -x (agent-local), 5 +turn, 6 -turn
    const float lx = (act == 3) ? dv : ((act == 4) ? -dv : 0.f);
    const float ly = (act == 1) ? dv : ((act == 2) ? -dv : 0.f);
    const float lw = (act == 5) ? dw : ((act == 6) ? -dw : 0.f);
    const float rad = __fmul_rn(0.017453292519943295f, ang);
    const float c = cosf(rad), s = sinf(rad);
    av = __fadd_rn(__fmul_rn(keep, av), lw);
    vel.x = __fadd_rn(__fmul_rn(keep, vel.x), __fsub_rn(__fmul_rn(c, lx), __fmul_rn(s, ly)));
    vel.y = __fadd_rn(__fmul_rn(keep, vel.y), __fadd_rn(__fmul_rn(s, lx), __fmul_rn(c, ly)));
}

// physics() as one kernel: one CTA per env, one warp per agent, the table read straight from HBM / L2 (an agent only
// touches the one or two runs around it, so staging the whole table would cost more than it saves).
__global__ void __launch_bounds__(128) physics_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    griddep_launch();        // view_kernel's CTAs may start staging their tables (static data) while this grid drains
    const int n = blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float* st_in = reinterpret_cast<float*>(smem_raw);
    float* st_out = st_in + A * ST_STRIDE;
    for (int a = tid; a < A; a += blockDim.x) {
        const int64_t i = (int64_t)n * A + a;
        const float2 pos = reinterpret_cast<const float2*>(k.a.positions)[i];
        float2 vel = reinterpret_cast<const float2*>(k.a.velocity)[i];
        const float ang = k.a.angles[i];
        float av = k.a.angvelocity[i];
        if (k.has_mv) momentum_movement(k, k.mv.actions[i], ang, av, vel);
        float* st = st_in + a * ST_STRIDE;
        st[ST_ANG] = ang; st[ST_PX] = pos.x; st[ST_PY] = pos.y; st[ST_AV] = av; st[ST_VX] = vel.x; st[ST_VY] = vel.y;
    }
    const int W = __ldg(k.s.line_widths + n) - AF;
    const int nb = (W + VRUN - 1) / VRUN;
    const int64_t b0 = __ldg(k.s.box_starts + n);
    const float4* occ = reinterpret_cast<const float4*>(k.s.occ_lines) + VRUN * b0;
    const float4* boxes = reinterpret_cast<const float4*>(k.s.occ_boxes) + b0;
    // msb_step: view_kernel is next and will stage this env's whole table — start it on its way from HBM to L2 now
    if (k.prefetch_view && tid == 0 && nb > 0) {
        bulk_prefetch_l2(occ, (uint32_t)nb * VRUN * 16u);
        bulk_prefetch_l2(k.s.occ_rec + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u);
        bulk_prefetch_l2(boxes, (uint32_t)nb * 16u);
    }
    const float rF = rcp(k.p.fps);
    const float r2 = fmul(k.p.agent_radius, 2.0020000934600830078f);
    const float r1 = fmul(k.p.agent_radius, 1.0010000467300415039f);
    __syncthreads();
    for (int a = warp; a < A; a += nwarps) {
        const float x = physics_agent<false>(st_in, A, a, lane, occ, boxes, W, nb, rF, r1, r2);
        if (lane == 0) physics_integrate(k, st_in, st_out, n, a, x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// lighting helpers
// ---------------------------------------------------------------------------------------------------------------

struct LaneLight { float x, y, i; int occ; };

// light_intensity() (kernels.cu:238-268) for a ray that hit an agent, evaluated by a whole warp with the first 32
// lights resident one per lane. Each lane remembers the static line that last occluded its light: consecutive
// agent-hit rays land centimetres apart, so one test per light (all lights in parallel) settles almost every
// occluded light; only the remaining lights are scanned against all static lines (lanes stride the lines, stop at
// the first occluder). Unoccluded lights are then accumulated in light order, exactly as the reference sums them.
// Which lines get tested varies; the occluded/unoccluded answer per light — hence the result — does not.
template <bool STATS>
__device__ __forceinline__ float light_intensity_cached(const float4* __restrict__ seg, int L, int AF, int I,
                                                        const float* lt, float Cx, float Cy, int lane, LaneLight& ll,
                                                        unsigned& iters) {
    const int nres = I < 32 ? I : 32;
    // phase 1: the remembered occluder of each resident light
    bool ob = false;
    if (lane < nres && ll.occ >= 0) {
        const Hit h = intersect(ll.x, ll.y, fsub(Cx, ll.x), fsub(Cy, ll.y), seg[ll.occ]);
        ob = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
    }
    const unsigned resident = nres == 32 ? 0xffffffffu : ((1u << nres) - 1u);
    unsigned todo = resident & ~__ballot_sync(0xffffffffu, ob);
    unsigned lit = 0;
    if (STATS) iters++;
    // phase 2: full scans for the lights the cache did not settle
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const float Ix = __shfl_sync(0xffffffffu, ll.x, i), Iy = __shfl_sync(0xffffffffu, ll.y, i);
        const float Ux = fsub(Cx, Ix), Uy = fsub(Cy, Iy);
        int found = -1;
        for (int base = AF; base < L; base += 64) {
            const int l0 = base + lane, l1 = base + 32 + lane;
            bool o0 = false, o1 = false;
            if (l0 < L) {
                const Hit h = intersect(Ix, Iy, Ux, Uy, seg[l0]);
                o0 = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
            }
            if (l1 < L) {
                const Hit h = intersect(Ix, Iy, Ux, Uy, seg[l1]);
                o1 = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
            }
            if (STATS) iters++;
            const unsigned b0 = __ballot_sync(0xffffffffu, o0), b1 = __ballot_sync(0xffffffffu, o1);
            if (b0 | b1) { found = b0 ? base + __ffs(b0) - 1 : base + 32 + __ffs(b1) - 1; break; }
        }
        if (found < 0) lit |= 1u << i;
        else if (lane == i) ll.occ = found;
    }
    // phase 3: sum the unoccluded lights in light order (:261-264)
    float acc = 0.1f;   // AMBIENT (kernels.cu:9)
    while (lit) {
        const int i = __ffs(lit) - 1;
        lit &= lit - 1;
        const float Ix = __shfl_sync(0xffffffffu, ll.x, i), Iy = __shfl_sync(0xffffffffu, ll.y, i);
        const float Ii = __shfl_sync(0xffffffffu, ll.i, i);
        const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
        acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);   // LUMINANCE = 2 (:240)
    }
    // lights beyond the first 32 (rare): the plain cooperative scan, still in light order
    for (int i = 32; i < I; i++) {
        const float Ix = __ldg(lt + 3 * i), Iy = __ldg(lt + 3 * i + 1), Ii = __ldg(lt + 3 * i + 2);
        const float Ux = fsub(Cx, Ix), Uy = fsub(Cy, Iy);
        bool occluded = false;
        for (int base = AF; base < L; base += 32) {
            const int l = base + lane;
            bool o = false;
            if (l < L) {
                const Hit h = intersect(Ix, Iy, Ux, Uy, seg[l]);
                o = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
            }
            if (__any_sync(0xffffffffu, o)) { occluded = true; break; }
        }
        if (!occluded) {
            const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
            acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);
        }
    }
    return fminf(acc, 1.f);
}

// light_intensity() evaluated by a single thread (bake: every lane has its own texel).
__device__ __forceinline__ float light_intensity_thread(const float4* seg, int AF, int L, const float* lt, int I,
                                                        float Cx, float Cy) {
    float acc = 0.1f;
    for (int i = 0; i < I; i++) {
        const float Ix = __ldg(lt + 3 * i), Iy = __ldg(lt + 3 * i + 1), Ii = __ldg(lt + 3 * i + 2);
        const float Ux = fsub(Cx, Ix), Uy = fsub(Cy, Iy);
        bool occluded = false;
        for (int l = AF; l < L; l++) {
            const Hit h = intersect(Ix, Iy, Ux, Uy, seg[l]);
            if ((h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f)) { occluded = true; break; }
        }
        if (!occluded) {
            const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
            acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);
        }
    }
    return fminf(acc, 1.f);
}

// ---------------------------------------------------------------------------------------------------------------
// view_kernel: render over the env's spatial table (static segments packed into runs of 16 with bounding boxes).
//
// One CTA per env, one warp per (agent, block of 32*NCH rays). Thread 0 stages the env's table — sorted segments,
// their line ids and the run boxes — with three 1-D bulk (TMA) copies on one mbarrier. Each warp then
//   1. puts one run box per lane, finds which of its 32-ray chunks each box can touch (half-plane tests against the
//      chunk boundaries in camera space) and a lower bound of its depth;
//   2. repeatedly takes the two NEAREST boxes that can still matter (REDUX min over a depth|lane key), bins their 32
//      segments (lane = segment: exact per-(agent, segment) terms of intersect(), chunk mask, depth bound), and runs
//      the reference's ray test, lane = ray, on the candidates a ballot yields per chunk. Front to back, boxes whose
//      depth bound lies behind the current hit of every ray they could touch are never visited: walls of the
//      agent's own room hide the rest of the floorplan (measured on the benchmark scenes: 4.5 of 18 runs visited);
//   3. the agents' model lines, last (mostly hidden by then).
// Exactness. The reference keeps, per ray, the hit with s < best - 1e-4 scanning lines in index order
// (kernels.cu:369-376): order-dependent. Here lines arrive in any order and a ray keeps its true minimum, flagging
// itself when two valid hits lie within AMB_EPS of each other. For an unflagged ray every other hit is more than
// 3e-4 behind the minimum, so the reference accepts the minimum when it reaches it and nothing after: same winner.
// A flagged ray (a ray through a wall corner: ~0.03% of rays) is replayed in line order by the whole warp, with the
// reference's rule. Culls only drop segments whose every hit is more than CULL_EPS (> AMB_EPS) behind the current
// minimum of every ray they could touch, so they change neither the minimum nor the flags' meaning.
// ---------------------------------------------------------------------------------------------------------------
constexpr float AMB_EPS = 3.e-4f;
constexpr float CULL_EPS = 4.e-4f;

struct VSmem {
    float4* seg;            // [AF + wcap]: [0, AF) the agents' model lines at their current poses; then the sorted static rows
    float4* boxes;          // [wcap / 16]
    int4* rec;              // [wcap] per sorted row: {texel offset lo, hi, texel count, line id} — only when KArgs::stage_rec
    const int4* rec_g;      // the same rows in global memory
    float4* scr;            // [nwarps][128] per warp: 64 candidate records while casting, then the chunk results
    float* st_in;           // [A][8]
    float* st_out;          // [A][8]
    int* mrad;              // [A] ([0]: bits of the model's radius, max |endpoint|)
    uint64_t* bar;
    int* next_item;         // items handed out beyond each warp's first
    int* meta;              // [16] this env's {W, lights, first light, first box, bits of occ_meta[0], [1], -, -} for queue entries;
                            //      {bits of vis_meta x0, y0, gx, gy, vis_starts lo, hi, -, -}
};

__device__ __forceinline__ VSmem vcarve(unsigned char* base, int wcap, int nwarps, int A, int AF, bool stage_rec) {
    VSmem m;
    m.seg = reinterpret_cast<float4*>(base);
    m.boxes = m.seg + AF + wcap;
    m.rec = reinterpret_cast<int4*>(m.boxes + wcap / VRUN);
    m.rec_g = nullptr;
    m.scr = reinterpret_cast<float4*>(m.rec + (stage_rec ? wcap : 0));
    m.st_in = reinterpret_cast<float*>(m.scr + nwarps * 128);
    m.st_out = m.st_in + A * ST_STRIDE;
    m.mrad = reinterpret_cast<int*>(m.st_out + A * ST_STRIDE);
    uintptr_t p = reinterpret_cast<uintptr_t>(m.mrad + A);
    p = (p + 15) & ~uintptr_t(15);
    m.bar = reinterpret_cast<uint64_t*>(p);
    m.next_item = reinterpret_cast<int*>(p + 8);
    m.meta = reinterpret_cast<int*>(p + 16);
    return m;
}

static size_t vsmem_bytes(int wcap, int nwarps, int A, int AF, bool stage_rec) {
    size_t b = (size_t)(AF + wcap) * 16 + (size_t)(wcap / VRUN) * 16 + (size_t)nwarps * 128 * 16 + (stage_rec ? (size_t)wcap * 16 : 0) +
               (size_t)A * ST_STRIDE * 4 * 2 + (size_t)A * 4;
    b = (b + 15) & ~size_t(15);
    return b + 16 + 64;
}

template <int NCH>
struct Rays {
    float rux[NCH], ruy[NCH], nearp[NCH], best[NCH], loc[NCH], cmax[NCH];
    float tie[NCH];         // the running minimum right after the latest near-tie (two hits within AMB_EPS), else +inf
    int row[NCH];           // winner's row in VSmem::seg; -1 = no hit
};

struct View { float px, py, cs, sn, xclip, B0, dB; };   // chunk c spans slopes (B0 - (c+1) dB, B0 - c dB) in camera space

// One batch of up to 32 segments against this warp's rays. Lane = segment while binning, lane = ray while testing.
template <int NCH, bool STATS>
__device__ __forceinline__ void cast_batch(const View& v, Rays<NCH>& ry, bool valid, float4 s4, int row,
                                           float4* __restrict__ scr, int lane, unsigned& tests) {
    // exact, ray-independent terms of intersect() (kernels.cu:83-85)
    const float Vx = fsub(s4.z, s4.x), Vy = fsub(s4.w, s4.y);
    const float PQx = fsub(s4.x, v.px), PQy = fsub(s4.y, v.py);
    scr[lane] = make_float4(Vx, Vy, PQx, PQy);
    *reinterpret_cast<float2*>(scr + 32 + lane) = make_float2(cross2(Vy, PQx, Vx, PQy), __int_as_float(row));
    // conservative summary in camera space (x' forward = the hit parameter s, y' left): culling only
    const float bxr = s4.z - v.px, byr = s4.w - v.py;
    const float xa = PQx * v.cs + PQy * v.sn, ya = PQy * v.cs - PQx * v.sn;
    const float xb = bxr * v.cs + byr * v.sn, yb = byr * v.cs - bxr * v.sn;
    unsigned cm = 0;
    if (valid && !((xa < v.xclip) && (xb < v.xclip))) {
        float ea = ya - xa * v.B0, eb = yb - xb * v.B0;             // > 0: left of the chunk's left boundary
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const float Bn = v.B0 - (float)(c + 1) * v.dB;
            const float ea1 = ya - xa * Bn, eb1 = yb - xb * Bn;     // < 0: right of the chunk's right boundary
            const bool outside = ((ea > 0.f) && (eb > 0.f)) || ((ea1 < 0.f) && (eb1 < 0.f));
            if (!outside) cm |= 1u << c;                            // (NaNs compare false: the exact test decides)
            ea = ea1; eb = eb1;
        }
    }
    const float smin = fminf(xa, xb) - 1e-3f - 1e-4f * fmaxf(fabsf(xa), fabsf(xb));
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        unsigned mask = __ballot_sync(0xffffffffu, ((cm >> c) & 1u) && !(smin > ry.cmax[c] + CULL_EPS));
        if (mask) {
            do {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 q = scr[j];
                const float2 w = *reinterpret_cast<const float2*>(scr + 32 + j);
                // raycast_kernel's test (kernels.cu:353-376), the reference's arithmetic op for op
                const float UxV = cross2(ry.rux[c], q.y, ry.ruy[c], q.x);
                const float rc = rcp(UxV);
                const float hs_ = fmul(w.x, rc);
                const float ht_ = fmul(cross2(ry.ruy[c], q.z, ry.rux[c], q.w), rc);
                const bool hit = !(fabsf(UxV) < PARALLEL_EPS) && (ht_ >= 0.f) && (ht_ <= 1.f) && (ry.nearp[c] < hs_);
                // near-tie bookkeeping: remember the minimum whenever a hit lands within AMB_EPS of the running minimum;
                // at the end the ray is ambiguous iff that memory is still within AMB_EPS of the final minimum
                const float d = hs_ - ry.best[c];
                const bool tie = hit && (fabsf(d) <= AMB_EPS), take = hit && (d < 0.f);
                ry.tie[c] = tie ? fminf(hs_, ry.best[c]) : ry.tie[c];
                ry.best[c] = take ? hs_ : ry.best[c];
                ry.loc[c] = take ? ht_ : ry.loc[c];
                ry.row[c] = take ? __float_as_int(w.y) : ry.row[c];
                if (STATS) tests++;
            } while (mask);
            // the chunk's farthest current hit; best >= 0 (or +inf), so its bit pattern orders like an unsigned integer
            ry.cmax[c] = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(ry.best[c])));
        }
    }
    __syncwarp();
}

// A flagged ray, replayed by the whole warp exactly as the reference scans it: lines in index order, lane = line,
// valid hits taken in ascending order with the s < best - 1e-4 rule (kernels.cu:353-376). Returns {s, t, line, dot}
// of the winner in every lane.
__device__ __noinline__ float4 replay_ray(const KArgs& k, const float4* __restrict__ dynseg, int64_t g0, int L, int AF,
                                             float px, float py, float ux, float uy, float nearp, float rlen, int lane) {
    const float4* __restrict__ lines = reinterpret_cast<const float4*>(k.s.lines) + g0;
    float best = CUDART_INF_F, bestm = CUDART_INF_F, loc = __int_as_float(0x7fffffff), dotv = __int_as_float(0x7fffffff);
    int idx = -1;
    for (int base = 0; base < L; base += 32) {
        const int l = base + lane;
        float hs_ = 0.f, ht_ = 0.f, Vx = 0.f, Vy = 0.f;
        bool hit = false;
        if (l < L) {
            const float4 s4 = l < AF ? dynseg[l] : __ldg(lines + l);
            Vx = fsub(s4.z, s4.x); Vy = fsub(s4.w, s4.y);
            const float PQx = fsub(s4.x, px), PQy = fsub(s4.y, py);
            const float UxV = cross2(ux, Vy, uy, Vx);
            const float rc = rcp(UxV);
            hs_ = fmul(cross2(Vy, PQx, Vx, PQy), rc);
            ht_ = fmul(cross2(uy, PQx, ux, PQy), rc);
            hit = !(fabsf(UxV) < PARALLEL_EPS) && (ht_ >= 0.f) && (ht_ <= 1.f) && (nearp < hs_);
        }
        unsigned hm = __ballot_sync(0xffffffffu, hit);
        while (hm) {
            const int i = __ffs(hm) - 1;
            hm &= hm - 1;
            const float si = __shfl_sync(0xffffffffu, hs_, i), ti = __shfl_sync(0xffffffffu, ht_, i);
            const float vxi = __shfl_sync(0xffffffffu, Vx, i), vyi = __shfl_sync(0xffffffffu, Vy, i);
            if (si < bestm) {
                best = si; bestm = fadd(si, -1.e-4f); loc = ti; idx = base + i;
                dotv = fmul(dot2(ux, vxi, uy, vyi), rcp(ffma(rlen, sqrt_(ffma(vxi, vxi, fmul(vyi, vyi))), 1.e-6f)));
            }
        }
    }
    return make_float4(best, loc, __int_as_float(idx), dotv);
}

// Dynamic light for the agent-hit rays of one chunk when they cannot be queued for dyn_kernel (no workspace, or the
// queue is full): the warp resolves them one ray at a time. Cold path, kept out of line.
__device__ __noinline__ float dyn_inline(const float4* seg, int L, int AF, int nlights, const float* lt, unsigned dm,
                                         float Cx, float Cy, float intensity) {
    const int lane = threadIdx.x & 31;
    LaneLight ll;
    ll.x = ll.y = ll.i = 0.f;
    ll.occ = -1;
    if (lane < nlights) { ll.x = __ldg(lt + 3 * lane); ll.y = __ldg(lt + 3 * lane + 1); ll.i = __ldg(lt + 3 * lane + 2); }
    unsigned iters = 0;
    while (dm) {
        const int j = __ffs(dm) - 1;
        dm &= dm - 1;
        const float cx = __shfl_sync(0xffffffffu, Cx, j), cy = __shfl_sync(0xffffffffu, Cy, j);
        const float v = light_intensity_cached<false>(seg, L, AF, nlights, lt, cx, cy, lane, ll, iters);
        if (lane == j) intensity = v;
    }
    return intensity;
}

// The sum over each group of `sub` adjacent lanes (sub a power of two; every lane of the group gets it), added as a
// butterfly from the widest stride down — (x0 + x2) + (x1 + x3) for sub = 4 — which is the order ATen's mean() adds the
// `sub` values of a pooled pixel in (probed: scripts/mean_order_probe.py), so that the fused RGB / Depth heads equal
// `downsample(x, sub).mean(-1)` (modules.py:181-183, 222-223) bit for bit (tests/test_gpu_reference_python.py).
__device__ __forceinline__ float pool_sum(float x, int sub, int lane) {
    (void)lane;
    for (int o = sub >> 1; o > 0; o >>= 1) x = __fadd_rn(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// What shading gathers for one ray: the two texels (and baked lights) around the hit, with the filter weights.
struct Texels { float lw, rw, tl0, tl1, tl2, tr0, tr1, tr2, bl, br; };

// shader_kernel (kernels.cu:407-450) + the Depth / RGB heads for one 32-ray chunk, lane = ray, in three steps so that
// several chunks' gathers can be in flight: shade_prepare (decode the parked hit, texel offset / count of its line from
// the staged table), shade_fetch (filter + texel / baked-light gathers; only issues loads), shade_chunk (the rest).
enum { ROW_UNKNOWN = 0x7fff };
struct ShadeIn { int l0; float locv, dotv, dist; int w; int64_t ts; bool hitany; };

__device__ __forceinline__ ShadeIn shade_prepare(const KArgs& k, const VSmem& m, int64_t g0, int AF, int r, float4 hitrec) {
    ShadeIn in;
    const int packed = __float_as_int(hitrec.x);
    in.l0 = packed < 0 ? -1 : (packed & 0xffff);
    in.locv = hitrec.y; in.dotv = hitrec.z; in.dist = hitrec.w;
    in.hitany = (r < k.p.res) && (packed >= 0);
    in.w = 0; in.ts = 0;
    if (in.hitany) {
        const int row = packed >> 16;
        if (row >= AF && row != ROW_UNKNOWN) {
            const int4 rc = k.stage_rec ? m.rec[row - AF] : __ldg(m.rec_g + (row - AF));
            in.l0 = rc.w;
            in.w = rc.z;
            in.ts = (int64_t)(((uint64_t)(uint32_t)rc.y << 32) | (uint32_t)rc.x);
        } else {                                            // an agent's model line, or a replayed ray: rare
            in.w = __ldg(k.s.tex_widths + g0 + in.l0);
            in.ts = __ldg(k.s.tex_starts + g0 + in.l0);
        }
    }
    return in;
}

// filter() (kernels.cu:394-405) + the gathers of shader_kernel (:427-430, :438)
__device__ __forceinline__ Texels shade_fetch(const KArgs& k, bool hitany, bool is_static, float locv, int w, int64_t ts) {
    Texels t;
    t.lw = t.rw = t.tl0 = t.tl1 = t.tl2 = t.tr0 = t.tr1 = t.tr2 = t.bl = t.br = 0.f;
    if (hitany) {
        const float yy = fminf(fmul(locv, (float)(w + 1)), (float)(w - 1));
        const int fl = __float2int_rz(fmaxf(fadd(yy, -1.f), 0.f));
        const int fr = __float2int_rz(yy);
        const float ld = fadd(fabsf(fsub(yy, (float)(fl + 1))), 1.e-3f);
        const float rd = fadd(fabsf(fsub(yy, (float)(fr + 1))), 1.e-3f);
        const float rc = rcp(fadd(rd, ld));
        t.lw = fmul(rd, rc);
        t.rw = fmul(ld, rc);
        const float* tl = k.s.textures + 3 * (ts + fl);
        const float* tr = k.s.textures + 3 * (ts + fr);
        t.tl0 = __ldg(tl); t.tl1 = __ldg(tl + 1); t.tl2 = __ldg(tl + 2);
        t.tr0 = __ldg(tr); t.tr1 = __ldg(tr + 1); t.tr2 = __ldg(tr + 2);
        if (is_static) { t.bl = __ldg(k.s.baked + ts + fl); t.br = __ldg(k.s.baked + ts + fr); }
    }
    return t;
}

__device__ __forceinline__ void shade_chunk(const KArgs& k, const VSmem& m, int n, int a, int AF, int Lrows,
                                            int r, int lane, const ShadeIn& in, const Texels& t) {
    const float4* __restrict__ seg = m.seg;
    const int A = k.s.n_agents, R = k.p.res;
    const int sub_ = k.has_obs ? k.obs.subsample : 1;
    const bool live = r < R;
    const int l0 = in.l0;
    const float locv = in.locv, dotv = in.dotv, dist = in.dist;
    const bool hitany = in.hitany;
    float intensity = 0.f, Cx = 0.f, Cy = 0.f;
    float kk0 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
    const bool isdyn = hitany && (l0 < AF);
    if (hitany) {
        if (!isdyn) {
            intensity = ffma(t.lw, t.bl, fmul(t.rw, t.br));                                           // :438
        } else {
            const float om = fsub(1.f, locv);                                                         // :435
            const float4 s4 = seg[l0];
            Cx = ffma(s4.x, om, fmul(locv, s4.z));
            Cy = ffma(s4.y, om, fmul(locv, s4.w));
        }
        kk0 = ffma(-dotv, dotv, 1.f);                                                                 // :442-445
        b0 = ffma(t.lw, t.tl0, fmul(t.rw, t.tr0));
        b1 = ffma(t.lw, t.tl1, fmul(t.rw, t.tr1));
        b2 = ffma(t.lw, t.tl2, fmul(t.rw, t.tr2));
    }
    // rays that hit an agent's model need the light at the hit point (:434-436): queue the chunk for dyn_kernel
    unsigned dm = __ballot_sync(0xffffffffu, isdyn);
    if (k.debug_skip_dyn) dm = 0;
    const int gl = lane & ~(sub_ - 1);                                    // first lane of my pixel group
    bool queued = false, deferred = false;
    if (dm) {
        const unsigned subm = sub_ == 32 ? 0xffffffffu : ((1u << sub_) - 1u);
        const unsigned gmask = (dm >> gl) & subm;                         // my group's agent-hit pixels
        if (k.dyn_entries) {
            // one entry per window of PS adjacent pixels that contains an agent-hit pixel (PS a multiple of sub_)
            const int PS = k.dyn_window, wl = lane & ~(PS - 1);
            const unsigned wmask = (dm >> wl) & (PS == 32 ? 0xffffffffu : ((1u << PS) - 1u));
            const unsigned leaders = __ballot_sync(0xffffffffu, wmask != 0 && lane == wl);
            const int cnt = __popc(leaders);
            int base = 0;
            if (lane == 0) base = atomicAdd(k.dyn_ctrl, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            queued = base + cnt <= k.dyn_cap;
            // which agent the window's first agent-hit pixel landed on: keys the persistent occluder cache
            const int tgt = __shfl_sync(0xffffffffu, l0, wl + (wmask ? __ffs(wmask) - 1 : 0)) / k.s.n_model;
            const int slot = base + __popc(leaders & ((1u << wl) - 1u));
            const bool mine = wmask != 0 && slot < k.dyn_cap;            // my window has a slot of the queue
            unsigned char* e = k.dyn_entries + (size_t)(mine ? slot : 0) * (DYN_HDR + 32 * PS);
            if (mine) {
                {
                    // a chunk that does not fit as a whole falls back inline: its reserved slots carry an empty mask
                    if (lane == wl) {
                        int4* h = reinterpret_cast<int4*>(e);
                        h[0] = make_int4(n, a * R + (r - lane + wl), queued ? (int)wmask : 0, sub_ | (tgt << 8));
                        h[1] = *reinterpret_cast<const int4*>(m.meta);          // what dyn_kernel would otherwise look up by env
                        *reinterpret_cast<int2*>(e + 32) = *reinterpret_cast<const int2*>(m.meta + 4);
                        // dyn_kernel's first loads: this (env, hit agent)'s occluder hints; the other lanes of the window
                        // below: the env's lights. Warm L2 now, a kernel ahead of their use.
                        if (queued) prefetch_l2(k.dyn_cache + ((size_t)n * A + (tgt < A ? tgt : 0)) * 32);
                    } else if (queued && lane - wl <= 3) {
                        const int line = lane - wl - 1;                           // 128-byte line of the env's lights (32 x 12 bytes: 3 lines)
                        if (32 * line < 3 * m.meta[1]) prefetch_l2(k.s.lights + 3 * (int64_t)m.meta[2] + 32 * line);
                    }
                    if (queued && gmask && live) {
                        // the lights that are certainly unoccluded from the hit point's cell of the visibility grid
                        unsigned sure = 0;
                        if (isdyn && !k.debug_no_vis) {
                            const int ix = __float2int_rd((Cx - __int_as_float(m.meta[8])) * VIS_INV_CELL);
                            const int iy = __float2int_rd((Cy - __int_as_float(m.meta[9])) * VIS_INV_CELL);
                            const int gx = m.meta[10], gy = m.meta[11];
                            if (ix >= 0 && ix < gx && iy >= 0 && iy < gy) {
                                const int64_t vs = (int64_t)(((uint64_t)(uint32_t)m.meta[13] << 32) | (uint32_t)m.meta[12]);
                                sure = __ldg(k.s.vis + vs + (int64_t)iy * gx + ix);
                            }
                        }
                        float4* rec = reinterpret_cast<float4*>(e + DYN_HDR) + 2 * (lane - wl);
                        rec[0] = make_float4(b0, b1, b2, kk0);
                        rec[1] = make_float4(Cx, Cy, intensity, __uint_as_float(sure));
                    }
                }
            }
            // publish: dyn_kernel may already be running (it is launched as a programmatic dependent and consumes the
            // queue while this grid drains). The window's lanes have written; one of them raises the entry's flag.
            // (release at GPU scope by one lane, after a warp barrier: cumulative over the other lanes' stores; no
            // acquire side here — __threadfence() would also invalidate the SM's L1 under the other CTAs' texel gathers)
            __syncwarp();
            if (mine && lane == wl) st_release(reinterpret_cast<int*>(e + DYN_FLAG), 1);
        }
        if (!queued) intensity = dyn_inline(seg, Lrows, AF, m.meta[1], k.s.lights + 3 * (int64_t)m.meta[2], dm, Cx, Cy, intensity);
        deferred = queued && gmask != 0;                                  // dyn_kernel writes this group's screen / rgb
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (hitany) {
        const float kk = fmul(kk0, intensity);
        s0 = fmul(kk, b0);
        s1 = fmul(kk, b1);
        s2 = fmul(kk, b2);
    }
    const int64_t ag = (int64_t)n * A + a;                 // warp-uniform: the 64-bit part of every address below
    const int64_t o0 = ag * R;
    const int om = k.out_mask;
    const uint32_t o32 = (uint32_t)ag * (uint32_t)R + (uint32_t)r;      // the same index when everything fits 32 bits (k.idx32)
    if (live && k.idx32) {
        // (one IMAD.WIDE.U32 per store; with 64-bit indices the compiler re-derives the whole product at every store)
        if (om & OUT_INDICES) k.out.indices[o32] = l0;
        if (om & OUT_LOCATIONS) k.out.locations[o32] = locv;
        if (om & OUT_DOTS) k.out.dots[o32] = dotv;
        if (om & OUT_DISTANCES) k.out.distances[o32] = dist;
        if ((om & OUT_SCREEN) && !(queued && isdyn)) { float* sc = k.out.screen + 3u * o32; sc[0] = s0; sc[1] = s1; sc[2] = s2; }
    } else if (live) {
        if (om & OUT_INDICES) (k.out.indices + o0)[r] = l0;
        if (om & OUT_LOCATIONS) (k.out.locations + o0)[r] = locv;
        if (om & OUT_DOTS) (k.out.dots + o0)[r] = dotv;
        if (om & OUT_DISTANCES) (k.out.distances + o0)[r] = dist;
        if ((om & OUT_SCREEN) && !(queued && isdyn)) { float* sc = k.out.screen + 3 * o0 + 3 * r; sc[0] = s0; sc[1] = s1; sc[2] = s2; }
    }
    // fused observation heads: Depth (modules.py:181-183) and RGB (:222-223), mean over `subsample` pixels
    if (k.has_obs) {
        float d = 0.f;
        if (live) {
            const float z = __fmul_rn(__fsub_rn(dist, k.p.agent_radius), k.inv_max_depth);
            d = __fsub_rn(1.f, fminf(fmaxf(z, 0.f), 1.f));
        }
        const float v0 = pool_sum(s0, sub_, lane), v1 = pool_sum(s1, sub_, lane), v2 = pool_sum(s2, sub_, lane), v3 = pool_sum(d, sub_, lane);
        if (live && lane == gl) {
            const int Ro = R >> k.sub_shift, ro = r >> k.sub_shift;         // subsample is a power of two
            const float inv = k.inv_sub;
            if (k.idx32) {
                const uint32_t od = (uint32_t)ag * (uint32_t)Ro + (uint32_t)ro;
                if ((om & OUT_RGB) && !deferred) {
                    float* q = k.obs.rgb + (3u * (uint32_t)ag * (uint32_t)Ro + (uint32_t)ro);
                    q[0] = __fmul_rn(v0, inv); q[Ro] = __fmul_rn(v1, inv); q[2 * Ro] = __fmul_rn(v2, inv);
                }
                if (om & OUT_DEPTH) k.obs.depth[od] = __fmul_rn(v3, inv);
            } else {
                if ((om & OUT_RGB) && !deferred) {
                    float* q = k.obs.rgb + ag * 3 * Ro + ro;
                    q[0] = __fmul_rn(v0, inv); q[Ro] = __fmul_rn(v1, inv); q[2 * Ro] = __fmul_rn(v2, inv);
                }
                if (om & OUT_DEPTH) (k.obs.depth + ag * Ro)[ro] = __fmul_rn(v3, inv);
            }
        }
    }
}

template <int NCH, bool STATS>
__device__ __forceinline__ void view_agent(const KArgs& k, const VSmem& m, int n, int64_t g0, int L, int W, int nb, int a,
                                           int rb, float4* __restrict__ scr, int lane) {
    const int A = k.s.n_agents, AF = A * k.s.n_model, R = k.p.res;
    const float* st = m.st_out + a * ST_STRIDE;
    View v;
    v.px = st[ST_PX]; v.py = st[ST_PY];
    v.sn = st[ST_SN]; v.cs = st[ST_CS];
    v.xclip = k.xclip;
    const float Rf = (float)R;
    const float rcpR = rcp(Rf);
    const int r0 = rb * (32 * NCH);
    v.B0 = (Rf - (float)(2 * r0)) * k.p.half_screen * rcpR;          // half a ray spacing left of ray r0
    v.dB = 64.f * k.p.half_screen * rcpR;

    // ---- rays (kernels.cu:341-344, ray_y :234-236). Lane = ray within each of this warp's NCH 32-ray chunks.
    Rays<NCH> ry;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int r = r0 + 32 * c + lane;
        const float y = fmul(fmul(fadd(fsub(Rf, (float)(unsigned)(2 * r)), -1.f), k.p.half_screen), rcpR);
        ry.rux[c] = ffma(v.sn, -y, v.cs);
        ry.ruy[c] = ffma(v.cs, y, v.sn);
        const float rlen = sqrt_(ffma(ry.rux[c], ry.rux[c], fmul(ry.ruy[c], ry.ruy[c])));
        ry.nearp[c] = fmul(rcp(rlen), k.p.agent_radius);
        ry.best[c] = r < R ? CUDART_INF_F : 0.f;                    // rays beyond R never take a hit
        ry.cmax[c] = (r0 + 32 * c) < R ? CUDART_INF_F : 0.f;
        ry.loc[c] = __int_as_float(0x7fffffff);
        ry.tie[c] = CUDART_INF_F;
        ry.row[c] = -1;
    }
    unsigned tests = 0, groups = 0, replays = 0;

    // ---- static lines: run boxes, nearest first; then the agents' model lines (kernels.cu:297-318 drew them), after
    // the walls that hide most of them. One loop, so that the batch code exists once.
    int bb = 0, dyn_next = 0;                       // first box of the current round of 32; next model line
    bool fresh = true;                              // the round's boxes have not been summarised yet
    unsigned key = 0xffffffffu, bcm = 0;
    float bsmin = 0.f;
    while (true) {
        bool valid;
        int row;                                    // row of VSmem::seg this lane bins
        if (bb < nb) {
            if (fresh) {
                fresh = false;
                key = 0xffffffffu; bcm = 0; bsmin = 0.f;
                if (bb + lane < nb) {
                    const float4 bx = m.boxes[bb + lane];
                    const float dx0 = bx.x - v.px, dx1 = bx.z - v.px, dy0 = bx.y - v.py, dy1 = bx.w - v.py;
                    const float cx0 = dx0 * v.cs, cx1 = dx1 * v.cs, sx0 = dx0 * v.sn, sx1 = dx1 * v.sn;
                    const float cy0 = dy0 * v.cs, cy1 = dy1 * v.cs, sy0 = dy0 * v.sn, sy1 = dy1 * v.sn;
                    // corners (x0,y0) (x1,y0) (x1,y1) (x0,y1) in camera space
                    const float X0 = cx0 + sy0, X1 = cx1 + sy0, X2 = cx1 + sy1, X3 = cx0 + sy1;
                    const float Y0 = cy0 - sx0, Y1 = cy0 - sx1, Y2 = cy1 - sx1, Y3 = cy1 - sx0;
                    const float xmax = fmaxf(fmaxf(X0, X1), fmaxf(X2, X3)), xmin = fminf(fminf(X0, X1), fminf(X2, X3));
                    if (!(xmax < v.xclip)) {
                        float B = v.B0;
                        bool left = (Y0 - X0 * B > 0.f) && (Y1 - X1 * B > 0.f) && (Y2 - X2 * B > 0.f) && (Y3 - X3 * B > 0.f);
#pragma unroll
                        for (int c = 0; c < NCH; c++) {
                            B = v.B0 - (float)(c + 1) * v.dB;
                            const float e0 = Y0 - X0 * B, e1 = Y1 - X1 * B, e2 = Y2 - X2 * B, e3 = Y3 - X3 * B;
                            const bool right = (e0 < 0.f) && (e1 < 0.f) && (e2 < 0.f) && (e3 < 0.f);
                            if (!left && !right) bcm |= 1u << c;
                            left = (e0 > 0.f) && (e1 > 0.f) && (e2 > 0.f) && (e3 > 0.f);
                        }
                        bsmin = fmaxf(xmin - 1e-3f - 1e-4f * fmaxf(fabsf(xmin), fabsf(xmax)), 0.f);
                        if (bcm) key = (__float_as_uint(bsmin) & ~31u) | (unsigned)lane;
                    }
                }
            }
            // a box is still worth visiting while some chunk it touches has a ray whose hit is not nearer than the box
            bool alive = false;
#pragma unroll
            for (int c = 0; c < NCH; c++) alive = alive || (((bcm >> c) & 1u) && !(bsmin > ry.cmax[c] + CULL_EPS));
            alive = alive && key != 0xffffffffu;
            const unsigned k0 = __reduce_min_sync(0xffffffffu, alive ? key : 0xffffffffu);
            if (k0 == 0xffffffffu) { bb += 32; fresh = true; continue; }
            const int j0 = (int)(k0 & 31u);
            const unsigned k1 = __reduce_min_sync(0xffffffffu, (alive && lane != j0) ? key : 0xffffffffu);
            const int j1 = k1 == 0xffffffffu ? -1 : (int)(k1 & 31u);
            if (lane == j0 || lane == j1) key = 0xffffffffu;
            const int run = lane < VRUN ? j0 : j1;
            const int srow = (bb + run) * VRUN + (lane & (VRUN - 1));
            valid = run >= 0 && srow < W;
            row = AF + (valid ? srow : 0);
        } else {
            if (dyn_next >= AF) break;
            if (dyn_next == 0) {
                // Can any agent's model be seen at all? Its lines lie within mrad of its position: if that disc is inside
                // the near plane's cull depth, wholly left / right of this warp's rays, or behind what every ray already
                // hit, cast_batch would drop each of its lines (same tests, per line) — and the own model always fails
                // the near-plane test (kernels.cu:369: its hits are nearer than agent_radius) when mrad < agent_radius.
                const float mrad = __int_as_float(m.mrad[0]);
                const float rho = mrad * 1.001f + 1e-3f;
                float cmaxall = ry.cmax[0];
#pragma unroll
                for (int c = 1; c < NCH; c++) cmaxall = fmaxf(cmaxall, ry.cmax[c]);
                const float Bn = v.B0 - (float)NCH * v.dB;
                const float sec0 = 1.0001f * sqrt_(1.f + v.B0 * v.B0), secn = 1.0001f * sqrt_(1.f + Bn * Bn);   // (rounded up: culling only)
                bool vis = false;
                for (int a2 = lane; a2 < A; a2 += 32) {
                    if (a2 == a) { vis = vis || !(mrad * 1.001f + 1e-4f < k.p.agent_radius); continue; }
                    const float* o = m.st_out + a2 * ST_STRIDE;
                    const float dx = o[ST_PX] - v.px, dy = o[ST_PY] - v.py;
                    const float X = dx * v.cs + dy * v.sn, Y = dy * v.cs - dx * v.sn;
                    const bool behind = X + rho < v.xclip;
                    const bool left = (Y - X * v.B0) - rho * sec0 > 0.f;
                    const bool right = (Y - X * Bn) + rho * secn < 0.f;
                    const bool hidden = X - rho - 1e-3f - 1e-4f * (fabsf(X) + rho) > cmaxall + CULL_EPS;
                    vis = vis || !(behind || left || right || hidden);      // (NaNs compare false: visible)
                }
                if (!__any_sync(0xffffffffu, vis)) break;
            }
            valid = dyn_next + lane < AF;
            row = valid ? dyn_next + lane : 0;
            dyn_next += 32;
        }
        cast_batch<NCH, STATS>(v, ry, valid, m.seg[row], row, scr, lane, tests);
        if (STATS) groups++;
    }

    // ---- per chunk: the winner's ray . line cosine (kernels.cu:362-364, winner only), near-tied rays replayed in line
    // order, results parked in shared memory for the shading loop as {row << 16 | line, location, dot, distance}
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int r = r0 + 32 * c + lane;
        const float rlen = sqrt_(ffma(ry.rux[c], ry.rux[c], fmul(ry.ruy[c], ry.ruy[c])));
        float dotv = __int_as_float(0x7fffffff);
        int packed = -1;
        if (ry.row[c] >= 0) {
            const int row = ry.row[c];
            const float4 s4 = m.seg[row];
            const float Vx = fsub(s4.z, s4.x), Vy = fsub(s4.w, s4.y);
            dotv = fmul(dot2(ry.rux[c], Vx, ry.ruy[c], Vy), rcp(ffma(rlen, sqrt_(ffma(Vx, Vx, fmul(Vy, Vy))), 1.e-6f)));
            packed = (row << 16) | (row < AF ? row : 0);              // a static row's line index comes with its record, at shading
        }
        float best = ry.best[c], loc = ry.loc[c];
        unsigned am = __ballot_sync(0xffffffffu, (ry.tie[c] - ry.best[c] <= AMB_EPS) && r < R);
        while (am) {
            const int j = __ffs(am) - 1;
            am &= am - 1;
            const float ux = __shfl_sync(0xffffffffu, ry.rux[c], j), uy = __shfl_sync(0xffffffffu, ry.ruy[c], j);
            const float np_ = __shfl_sync(0xffffffffu, ry.nearp[c], j), rl = __shfl_sync(0xffffffffu, rlen, j);
            const float4 w = replay_ray(k, m.seg, g0, L, AF, v.px, v.py, ux, uy, np_, rl, lane);
            if (lane == j) {
                best = w.x; loc = w.y; dotv = w.w;
                const int l = __float_as_int(w.z);                  // its row is unknown: shading looks its texels up by line
                packed = l < 0 ? -1 : ((l < AF ? l : ROW_UNKNOWN) << 16) | l;
            }
            if (STATS) replays++;
        }
        scr[32 * c + lane] = make_float4(__int_as_float(packed), loc, dotv, fmul(rlen, best));
    }
    __syncwarp();
    // ---- shading: the texel gathers of two chunks in flight at a time (all four cost more in spills than they hide)
#pragma unroll 1
    for (int c0 = 0; c0 < NCH; c0 += 2) {
        ShadeIn in[2];
        Texels tx[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (c0 + u < NCH) {
                in[u] = shade_prepare(k, m, g0, AF, r0 + 32 * (c0 + u) + lane, scr[32 * (c0 + u) + lane]);
                tx[u] = shade_fetch(k, in[u].hitany, in[u].l0 >= AF, in[u].locv, in[u].w, in[u].ts);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (c0 + u < NCH) shade_chunk(k, m, n, a, AF, AF + W, r0 + 32 * (c0 + u) + lane, lane, in[u], tx[u]);
        }
    }
    __syncwarp();
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_TESTS, (unsigned long long)tests);
        atomicAdd(k.stats + STAT_GROUPS, (unsigned long long)groups);
        atomicAdd(k.stats + STAT_REPLAYS, (unsigned long long)replays);
    }
}

template <int NCH, bool PHYS, bool STATS>
__global__ void __launch_bounds__(MSB_VIEW_THREADS, MSB_VIEW_BLOCKS) view_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // costliest envs first: a short tail. (Through REDUX so that n — and every address part derived from it — lives in
    // a uniform register.)
    const int n = k.env_order ? __reduce_max_sync(0xffffffffu, __ldg(k.env_order + blockIdx.x)) : (int)blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    VSmem m = vcarve(smem_raw, k.wcap, nwarps, A, AF, k.stage_rec != 0);
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    const int W = L - AF;
    const int nb = (W + VRUN - 1) / VRUN;
    m.rec_g = reinterpret_cast<const int4*>(k.s.occ_rec) + VRUN * (int64_t)__ldg(k.s.box_starts + n);
    // stage this env's table: two or three bulk (TMA) copies, ragged-packed HBM -> shared memory, one mbarrier
    if (tid == 32 % blockDim.x) {
        m.meta[0] = W; m.meta[1] = __ldg(k.s.light_widths + n); m.meta[2] = __ldg(k.s.light_starts + n);
        m.meta[3] = __ldg(k.s.box_starts + n);
        m.meta[4] = __float_as_int(__ldg(k.s.occ_meta + 2 * n)); m.meta[5] = __float_as_int(__ldg(k.s.occ_meta + 2 * n + 1));
        m.meta[6] = 0; m.meta[7] = 0;
        m.meta[10] = 0; m.meta[11] = 0;                                  // no grid: every lookup falls outside
        if (k.s.vis) {
            const float4 vm = __ldg(reinterpret_cast<const float4*>(k.s.vis_meta) + n);
            const int64_t vs = __ldg(k.s.vis_starts + n);
            m.meta[8] = __float_as_int(vm.x); m.meta[9] = __float_as_int(vm.y); m.meta[10] = (int)vm.z; m.meta[11] = (int)vm.w;
            m.meta[12] = (int)(uint32_t)vs; m.meta[13] = (int)(uint32_t)((uint64_t)vs >> 32);
        }
    }
    if (tid == 0) {
        m.mrad[0] = 0;
        *m.next_item = 0;
        mbar_init(m.bar, 1);
        if (nb > 0) {
            const int64_t b0 = __ldg(k.s.box_starts + n);
            mbar_expect_tx(m.bar, (uint32_t)nb * (VRUN * (k.stage_rec ? 32u : 16u) + 16u));
            bulk_g2s(m.seg + AF, k.s.occ_lines + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u, m.bar);
            if (k.stage_rec) bulk_g2s(m.rec, k.s.occ_rec + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u, m.bar);
            bulk_g2s(m.boxes, k.s.occ_boxes + 4 * b0, (uint32_t)nb * 16u, m.bar);
        }
    }
    // Everything above reads static scenery only. The agents' state (and everything written below) belongs to the
    // kernels ahead in the stream: wait for them here. dyn_kernel may queue up behind this grid from now on.
    griddep_wait();
    griddep_launch();
    // agent state -> shared memory (through MomentumMovement when the step carries actions)
    for (int a = tid; a < A; a += blockDim.x) {
        const int64_t i = (int64_t)n * A + a;
        const float2 pos = reinterpret_cast<const float2*>(k.a.positions)[i];
        float2 vel = reinterpret_cast<const float2*>(k.a.velocity)[i];
        const float ang = k.a.angles[i];
        float av = k.a.angvelocity[i];
        if (PHYS && k.has_mv) momentum_movement(k, k.mv.actions[i], ang, av, vel);
        float* st = (PHYS ? m.st_in : m.st_out) + a * ST_STRIDE;
        st[ST_ANG] = ang; st[ST_PX] = pos.x; st[ST_PY] = pos.y; st[ST_AV] = av; st[ST_VX] = vel.x; st[ST_VY] = vel.y;
        if (!PHYS) sincos_deg(ang, st[ST_SN], st[ST_CS]);       // once per agent (kernels.cu:304-306, 335-337 redo it per thread)
    }
    __syncthreads();
    if (nb > 0) mbar_wait(m.bar, 0);
    if (PHYS) {
        // physics (kernels.cu:179-230) over the staged table: one warp per agent, then the fused integration
        const float rF = rcp(k.p.fps);
        const float r2 = fmul(k.p.agent_radius, 2.0020000934600830078f);
        const float r1 = fmul(k.p.agent_radius, 1.0010000467300415039f);
        for (int a = warp; a < A; a += nwarps) {
            const float x = physics_agent<true>(m.st_in, A, a, lane, m.seg + AF, m.boxes, W, nb, rF, r1, r2);
            if (lane == 0) physics_integrate(k, m.st_in, m.st_out, n, a, x, true);
        }
        __syncthreads();
    }
    // draw_kernel (kernels.cu:297-318): the agents' model lines, at their current poses, into shared and global memory
    {
        const int F = k.s.n_model;
        for (int t = tid; t < 2 * F; t += blockDim.x) {         // the model's radius: lets a warp skip agents it cannot see
            const float2 mp = __ldg(reinterpret_cast<const float2*>(k.s.model) + t);
            atomicMax(m.mrad, __float_as_int(sqrtf(mp.x * mp.x + mp.y * mp.y)));
        }
        for (int a = warp; a < A; a += nwarps) {
            const float* st = m.st_out + a * ST_STRIDE;
            const float sn = st[ST_SN], cs = st[ST_CS];
            for (int t = lane; t < 2 * F; t += 32) {            // t = 2 * model line + endpoint
                const float2 mp = __ldg(reinterpret_cast<const float2*>(k.s.model) + t);
                const float2 pt = make_float2(fadd(st[ST_PX], cross2(cs, mp.x, sn, mp.y)), fadd(st[ST_PY], dot2(sn, mp.x, cs, mp.y)));
                reinterpret_cast<float2*>(m.seg)[2 * a * F + t] = pt;
                reinterpret_cast<float2*>(k.s.lines)[2 * (g0 + a * F) + t] = pt;
            }
        }
    }
    __syncthreads();
    // (agent, block of rays) items: the first is the warp's own, the following ones come off a shared counter — items
    // differ in cost by what they see, and the CTA holds its registers and shared memory until its last warp is done
    const int RB = k.ray_blocks, rbs = k.rb_shift;          // rb_shift >= 0: RB is that power of two
    for (int w = warp; w < A * RB;) {
        const int a = rbs >= 0 ? w >> rbs : w / RB;
        view_agent<NCH, STATS>(k, m, n, g0, L, W, nb, a, w - a * RB, m.scr + warp * 128, lane);
        int next = 0;
        if (lane == 0) next = nwarps + atomicAdd(m.next_item, 1);
        w = __reduce_max_sync(0xffffffffu, next);            // (lands in a uniform register: so do the item's address parts)
    }
    if (k.has_obs && k.obs.imu) {
        for (int a = tid; a < A; a += blockDim.x) {
            const float* st = m.st_out + a * ST_STRIDE;
            const float ang = __fmul_rn(0.017453292519943295f, st[ST_ANG]);
            const float c = cosf(ang), s = sinf(ang);
            const float vx = st[ST_VX], vy = st[ST_VY];
            float* q = k.obs.imu + 3 * ((int64_t)n * A + a);
            q[0] = __fmul_rn(st[ST_AV], k.inv_ang);
            q[1] = __fmul_rn(__fadd_rn(__fmul_rn(c, vx), __fmul_rn(s, vy)), k.inv_speed);
            q[2] = __fmul_rn(__fadd_rn(__fmul_rn(-s, vx), __fmul_rn(c, vy)), k.inv_speed);
        }
    }
    // tell dyn_kernel that this CTA will publish nothing more (its last act: dyn_kernel outlives the grid)
    if (k.dyn_entries) {
        __syncthreads();
        if (tid == 0) red_release_add(k.dyn_ctrl + 3, 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dyn_kernel: the load-balanced second pass over ray chunks that contain agent-hit pixels.
// Entry = a window of PS = max(4, subsample) adjacent pixels with at least one agent-hit pixel: 16-byte header {env,
// agent*R + first ray of the window, mask of agent-hit pixels, subsample | hit agent << 8} + per pixel two float4:
// {texel rgb, 1-dot^2} and {hit point x, hit point y, static intensity, -}; only the pixels of pooling groups that
// contain an agent-hit pixel are filled in. Small windows bound the serial work per entry; sharing the scans between
// the few pixels of a window (they land centimetres apart on one agent's outline) divides the work per pixel.
// One warp per entry, entries strided over one resident wave of warps (every entry is an independent unit of work, so the
// whole machine is busy whatever the scene). light_intensity() (kernels.cu:238-268) for all the entry's pixels at once:
//   1. lane = light: each light's remembered occluder (persistent per (env, hit agent) in the workspace; a hint,
//      re-verified by the exact test) settles most occluded lights, one test per (pixel, light);
//   2. light by light, for the lights some pixel still needs: lane b tests run b's box against the box around the
//      light and ALL the entry's hit points (they lie on one agent's outline, centimetres apart), grown by a margin
//      covering the worst-case rounding of intersect(); the overlapping runs are read once, two per iteration
//      (lane = segment), and tested against every pixel that still needs the light, until it is occluded;
//   3. lane = pixel: unoccluded lights are summed in light order, as the reference does.
// Which segments get tested varies; the occluded / unoccluded answer per (pixel, light) — hence the result — does not.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool occludes(const Hit h) { return (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f); }

// One queue entry as every lane of a warp holds it after dyn_scan: the header's fields (warp-uniform), this lane's pixel
// (lane < PS) and this lane's light (lane < nlights), plus what the scan found.
struct DynEntry {
    int n, sub, tgt, nlights, R0;      // env; pooling factor; hit agent; lights of the env (0: empty slot); agent * R + first ray
    unsigned mask;                     // the window's agent-hit pixels
    bool have, isdyn;                  // my pixel's records were filled in; my pixel hit an agent
    float4 ra, rb;                     // my pixel: {texel rgb, 1 - dot^2}, {hit point, static intensity, bits vouched for by the grid}
    float lx, ly, li;                  // my light
    unsigned mylit;                    // lights (of this warp's share) with nothing between them and my pixel
    float intensity;                   // my pixel's light when it does not come from `mylit` (static pixel; > 32 lights)
};

// Steps 1 and 2 of the second pass for one entry, by one warp that owns the lights i = part (mod parts). Leaves in
// d.mylit the owned lights found unoccluded per pixel lane; with more than 32 lights part 0 alone computes d.intensity.
template <bool STATS>
__device__ __forceinline__ void dyn_scan(const KArgs& k, const unsigned char* e, int part, int parts, int lane, DynEntry& d,
                                         unsigned& dyn_iters) {
    const int A = k.s.n_agents, AF = A * k.s.n_model, R = k.p.res;
    const int PS = k.dyn_window;
    // (entries are read past L1: a neighbour's line fetched earlier by this SM may hold this entry's bytes from
    // before they were written)
    const int4 hdr = __ldcg(reinterpret_cast<const int4*>(e));
    const int4 hdr1 = __ldcg(reinterpret_cast<const int4*>(e + 16));      // {W, lights, first light, first box} of the env
    const float2 hdr2 = __ldcg(reinterpret_cast<const float2*>(e + 32));   // occ_meta of the env
    const unsigned mask = (unsigned)hdr.z;                          // 0: slot reserved by a chunk that fell back inline
    const int sub = hdr.w & 0xff, tgt = hdr.w >> 8;
    const int n = hdr.x, ar = hdr.y;                       // env; agent * R + first ray of the window
    const int r0 = ar % R;
    const int gl = lane & ~(sub - 1);
    const unsigned subm = sub == 32 ? 0xffffffffu : ((1u << sub) - 1u);
    const unsigned gmask = lane < PS ? (mask >> gl) & subm : 0u;   // my pooling group's agent-hit pixels
    const bool have = lane < PS && gmask != 0 && (r0 + lane < R);       // my pixel's records were filled in
    const bool isdyn = lane < PS && ((mask >> lane) & 1u);
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) {
        const float4* rec = reinterpret_cast<const float4*>(e + DYN_HDR) + 2 * lane;
        ra = __ldcg(rec);
        rb = __ldcg(rec + 1);
    }
    const float Cx = rb.x, Cy = rb.y;
    const unsigned sure = (isdyn && !k.debug_no_vis) ? __float_as_uint(rb.w) : 0u;   // lights the visibility grid vouches for
    const int W = hdr1.x, L = W + AF, nb = (W + VRUN - 1) / VRUN;
    const int nlights = mask ? hdr1.y : 0;
    const float* lt = k.s.lights + 3 * (int64_t)hdr1.z;
    float intensity = rb.z;
    unsigned mylit = 0;
    float lx = 0.f, ly = 0.f, li = 0.f;
    if (nlights > 32) {
        // rare: more lights than lanes. Part 0 alone, one pixel at a time over the env's lines in their original order.
        if (part == 0) {
            const float4* seg = reinterpret_cast<const float4*>(k.s.lines) + __ldg(k.s.line_starts + n);
            LaneLight ll;
            ll.occ = -1;
            ll.x = __ldg(lt + 3 * lane); ll.y = __ldg(lt + 3 * lane + 1); ll.i = __ldg(lt + 3 * lane + 2);
            for (unsigned m = mask; m; m &= m - 1) {
                const int p = __ffs(m) - 1;
                const float cx = __shfl_sync(0xffffffffu, Cx, p), cy = __shfl_sync(0xffffffffu, Cy, p);
                const float v = light_intensity_cached<STATS>(seg, L, AF, nlights, lt, cx, cy, lane, ll, dyn_iters);
                if (lane == p) intensity = v;
            }
        }
    } else if (nlights > 0) {
        const int64_t b0 = hdr1.w;
        const float4* occ = reinterpret_cast<const float4*>(k.s.occ_lines) + VRUN * b0;
        const float4* boxes = reinterpret_cast<const float4*>(k.s.occ_boxes) + b0;
        const float vmax = hdr2.x, diam = hdr2.y;
        // lights one per lane, each with the occluder remembered for this (env, agent that was hit)
        int* cache = k.dyn_cache + ((size_t)n * A + (tgt < A ? tgt : 0)) * 32;
        int hint = cache[lane];
        if (lane < nlights) { lx = __ldg(lt + 3 * lane); ly = __ldg(lt + 3 * lane + 1); li = __ldg(lt + 3 * lane + 2); }
        const int hint_before = hint;
        const unsigned resident = nlights == 32 ? 0xffffffffu : ((1u << nlights) - 1u);
        float4 bx0 = make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        if (lane < nb) bx0 = __ldg(boxes + lane);
        // 1. the remembered occluders (every warp, redundantly: one test per pixel)
        const bool has_hint = lane < nlights && hint >= 0 && hint < W;
        float4 hseg = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_hint) hseg = __ldg(occ + hint);
        unsigned mytodo = 0;
        for (unsigned m = mask; m; m &= m - 1) {
            const int p = __ffs(m) - 1;
            const float cx = __shfl_sync(0xffffffffu, Cx, p), cy = __shfl_sync(0xffffffffu, Cy, p);
            bool ob = false;
            if (has_hint) ob = occludes(intersect(lx, ly, fsub(cx, lx), fsub(cy, ly), hseg));
            const unsigned todo = resident & ~__ballot_sync(0xffffffffu, ob);
            if (lane == p) mytodo = todo & ~sure;
        }
        if (STATS) dyn_iters++;
        // the box around the entry's hit points
        float cx0 = isdyn ? Cx : CUDART_INF_F, cx1 = isdyn ? Cx : -CUDART_INF_F;
        float cy0 = isdyn ? Cy : CUDART_INF_F, cy1 = isdyn ? Cy : -CUDART_INF_F;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx0 = fminf(cx0, __shfl_xor_sync(0xffffffffu, cx0, o)); cx1 = fmaxf(cx1, __shfl_xor_sync(0xffffffffu, cx1, o));
            cy0 = fminf(cy0, __shfl_xor_sync(0xffffffffu, cy0, o)); cy1 = fmaxf(cy1, __shfl_xor_sync(0xffffffffu, cy1, o));
        }
        // 2. scans, light by light; part w owns the lights i = w (mod parts). (Ownership must not depend on which lights
        // still need a scan: the hints are shared with other warps and may change between two warps' reads.)
        unsigned owned = 0x11111111u;                                   // parts = 4
        if (parts == 1) owned = 0xffffffffu; else if (parts == 2) owned = 0x55555555u; else if (parts == 3) owned = 0x49249249u;
        unsigned todo_any = __reduce_or_sync(0xffffffffu, isdyn ? mytodo : 0u) & (owned << part);
        const int slot = lane / VRUN, within = lane - slot * VRUN;
        while (todo_any) {
            const int i = __ffs(todo_any) - 1;
            todo_any &= todo_any - 1;
            const float Ix = __shfl_sync(0xffffffffu, lx, i), Iy = __shfl_sync(0xffffffffu, ly, i);
            unsigned need = __ballot_sync(0xffffffffu, isdyn && ((mytodo >> i) & 1u));
            const unsigned iters0 = dyn_iters;
            // conservative query (see DESIGN.md "shadow cull"): rounding can move the computed crossing by at most
            // delta (a fraction of each segment's length) along either segment. A run can only hold an occluder if
            // its box comes within mg (+ the spread of the hit points) of the segment from the light to the middle
            // of the hit points: slab test of that segment against the grown box.
            const float ulen = fmaxf(fmaxf(fabsf(cx0 - Ix), fabsf(cx1 - Ix)), fmaxf(fabsf(cy0 - Iy), fabsf(cy1 - Iy)));
            const float delta = 4e-4f * vmax * (diam + ulen);
            const float mg = delta * (ulen + vmax) + 0.01f + fmaxf(cx1 - cx0, cy1 - cy0);
            const float mx_ = 0.5f * (cx0 + cx1) - Ix, my_ = 0.5f * (cy0 + cy1) - Iy;      // light -> middle of the hit points
            const float irx = 1.f / mx_, iry = 1.f / my_;                                  // +-inf when axis-parallel
            int found = -1;
            for (int bb = 0; bb < nb && need; bb += 32) {
                float4 bx = bx0;
                if (bb) bx = (bb + lane < nb) ? __ldg(boxes + bb + lane) : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
                // parameters at which the segment is inside each slab of the grown box; NaN (0 * inf: the segment runs
                // along a slab boundary) compares false below, i.e. errs on the side of visiting
                const float tx0 = (bx.x - mg - Ix) * irx, tx1 = (bx.z + mg - Ix) * irx;
                const float ty0 = (bx.y - mg - Iy) * iry, ty1 = (bx.w + mg - Iy) * iry;
                const float tlo = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), 0.f);
                const float thi = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), 1.f);
                const bool visit = (bb + lane < nb) && !(tlo > thi);
                unsigned runs = __ballot_sync(0xffffffffu, visit);
                // lanes [0, 16) take the first run still to visit, lanes [16, 32) the second; the next pair's
                // segments are requested before this pair's are tested (one L2 round trip per light, not per pair)
                int l = W;
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (runs) {
                    const unsigned rest = runs & (runs - 1);
                    const int nth = slot == 0 ? __ffs(runs) - 1 : (rest ? __ffs(rest) - 1 : -1);
                    l = nth >= 0 ? VRUN * (bb + nth) + within : W;
                    if (l < W) s4 = __ldg(occ + l);
                }
                while (runs && need) {
                    const unsigned rest = runs & (runs - 1);
                    const unsigned after = rest & (rest - 1);
                    int l_next = W;
                    float4 s4_next = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (after) {
                        const unsigned rest2 = after & (after - 1);
                        const int nth = slot == 0 ? __ffs(after) - 1 : (rest2 ? __ffs(rest2) - 1 : -1);
                        l_next = nth >= 0 ? VRUN * (bb + nth) + within : W;
                        if (l_next < W) s4_next = __ldg(occ + l_next);
                    }
                    const bool real = l < W;
                    float Vx = 0.f, Vy = 0.f, PQx = 0.f, PQy = 0.f, snum = 0.f;
                    if (real) {
                        Vx = fsub(s4.z, s4.x); Vy = fsub(s4.w, s4.y);
                        PQx = fsub(s4.x, Ix); PQy = fsub(s4.y, Iy);
                        snum = cross2(Vy, PQx, Vx, PQy);
                    }
                    for (unsigned q = need; q; q &= q - 1) {
                        const int p = __ffs(q) - 1;
                        const float ux = fsub(__shfl_sync(0xffffffffu, Cx, p), Ix), uy = fsub(__shfl_sync(0xffffffffu, Cy, p), Iy);
                        const bool o = real && occludes(intersect_pre(ux, uy, Vx, Vy, PQx, PQy, snum));
                        const unsigned bal = __ballot_sync(0xffffffffu, o);
                        if (bal) { need &= ~(1u << p); found = __shfl_sync(0xffffffffu, l, __ffs(bal) - 1); }
                    }
                    if (STATS) dyn_iters++;
                    runs = after; l = l_next; s4 = s4_next;
                }
            }
            if (found >= 0 && lane == i) hint = found;
            if ((need >> lane) & 1u) mylit |= 1u << i;         // nothing in the way of light i for my pixel
            if (STATS && k.stats && lane == 0) {
                atomicAdd(k.stats + STAT_DYN_SCANS, 1ull);
                if (need) { atomicAdd(k.stats + STAT_DYN_SCANS_LIT, 1ull); atomicAdd(k.stats + STAT_DYN_ITERS_LIT, (unsigned long long)(dyn_iters - iters0)); }
            }
        }
        if (hint != hint_before) cache[lane] = hint;   // racy on purpose: any stored value is only a hint
    }
    d.n = n; d.sub = sub; d.tgt = tgt; d.nlights = nlights; d.R0 = ar; d.mask = mask;
    d.have = have; d.isdyn = isdyn; d.ra = ra; d.rb = rb; d.lx = lx; d.ly = ly; d.li = li;
    d.mylit = mylit; d.intensity = intensity;
}

// Step 3 of the second pass, lane = pixel: the unoccluded lights (`lit`: every part's finds plus what the visibility grid
// vouches for) summed in light order (kernels.cu:261-264), then the pixel's screen value and its pooled RGB.
__device__ __forceinline__ void dyn_finish(const KArgs& k, const DynEntry& d, unsigned lit_mine, int lane) {
    const int A = k.s.n_agents, R = k.p.res;
    const int av = d.R0 / R, r0 = d.R0 - av * R;
    const int64_t ag = (int64_t)d.n * A + av;
    const int sub = d.sub, gl = lane & ~(sub - 1);
    float intensity = d.intensity;
    if (d.nlights <= 32) {
        const float Cx = d.rb.x, Cy = d.rb.y;
        float acc = 0.1f;                                  // AMBIENT (kernels.cu:9)
        for (unsigned lit = __reduce_or_sync(0xffffffffu, lit_mine); lit; lit &= lit - 1) {
            const int i = __ffs(lit) - 1;
            const float Ix = __shfl_sync(0xffffffffu, d.lx, i), Iy = __shfl_sync(0xffffffffu, d.ly, i);
            const float Ii = __shfl_sync(0xffffffffu, d.li, i);
            if ((lit_mine >> i) & 1u) {
                const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
                acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);   // LUMINANCE = 2 (:240)
            }
        }
        if (d.isdyn) intensity = fminf(acc, 1.f);
    }
    const float kk = fmul(d.ra.w, intensity);
    const float s0 = fmul(kk, d.ra.x), s1 = fmul(kk, d.ra.y), s2 = fmul(kk, d.ra.z);
    if (k.out.screen && d.isdyn) {
        float* sc = k.out.screen + 3 * (ag * R + r0 + lane);
        sc[0] = s0; sc[1] = s1; sc[2] = s2;
    }
    if (k.has_obs && k.obs.rgb) {
        const float v0 = pool_sum(d.have ? s0 : 0.f, sub, lane), v1 = pool_sum(d.have ? s1 : 0.f, sub, lane), v2 = pool_sum(d.have ? s2 : 0.f, sub, lane);
        if (d.have && lane == gl) {
            const int Ro = R >> k.sub_shift, ro = (r0 + lane) >> k.sub_shift;
            float* q = k.obs.rgb + ag * 3 * Ro + ro;
            q[0] = __fmul_rn(v0, k.inv_sub); q[Ro] = __fmul_rn(v1, k.inv_sub); q[2 * Ro] = __fmul_rn(v2, k.inv_sub);
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(128, MSB_DYN_BLOCKS) dyn_kernel(const __grid_constant__ KArgs k) {
    __shared__ unsigned s_lit[4][32];                          // per warp: the lights it found unoccluded, per pixel lane
    __shared__ int s_next, s_ready;                            // the CTA's next entry; whether the current one was published
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    unsigned dyn_rays = 0, dyn_iters = 0, dyn_entries = 0;
    // One entry per CTA at a time (every entry is an independent unit of work). The CTA's warps split the entry's LIGHTS
    // between them: the kernel lasts as long as its slowest entry (an agent in a large open room: ~20 lights to scan one
    // after the other), so the per-entry chain is what has to be short.
    const long long t_start = STATS ? clock64() : 0;
    long long t_entries = 0;
    // the first entry is the CTA's own index; the following ones come off a shared counter (ctrl[2]) — CTAs that drew
    // heavy entries take fewer. The counter is bumped at the start of an entry, so its round trip is off the chain.
    // The queue is consumed WHILE view_kernel fills it: this grid is launched as a programmatic dependent, its CTAs
    // move in as view_kernel's leave (every one of those has started by then, so nothing here can starve them). An
    // entry is ready when its flag is up; it never will be once every CTA of view_kernel has signed off with the flag
    // still down — slots are reserved in order, so neither will any later one.
    for (int ei = blockIdx.x; ei < k.dyn_cap;) {
        const long long t_e0 = STATS ? clock64() : 0;
        const int PS = k.dyn_window;
        unsigned char* e = k.dyn_entries + (size_t)ei * (DYN_HDR + 32 * PS);
        int nxt = 0;
        if (threadIdx.x == 0) {
            volatile int* flag = reinterpret_cast<volatile int*>(e + DYN_FLAG);
            volatile int* done = k.dyn_ctrl + 3;
            int ready = *flag;
            bool over = false;
            for (unsigned spins = 0; !ready && !over && spins < (1u << 17); spins++) {
                if (*done >= k.view_ctas) { __threadfence(); ready = *flag; over = true; break; }
                __nanosleep(100);
                ready = *flag;
            }
            if (!ready && !over) {
                // polled for tens of milliseconds: stop guessing and block until view_kernel has completed as a grid
                griddep_wait();
                __threadfence();
                ready = *flag;
            }
            if (ready) {
                __threadfence();
                *flag = 0;                                                          // re-armed for the next step
                nxt = atomicAdd(k.dyn_ctrl + 2, 1) + (int)gridDim.x;
            }
            s_ready = ready;
        }
        __syncthreads();
        if (!s_ready) break;
        DynEntry d;
        dyn_scan<STATS>(k, e, warp, NW, lane, d, dyn_iters);
        if (STATS && warp == 0 && d.mask) { dyn_rays += __popc(d.mask); dyn_entries++; }
        s_lit[warp][lane] = d.mylit;
        if (threadIdx.x == 0) s_next = nxt;
        __syncthreads();
        const int ei_next = s_next;
        if (warp == 0) {
            unsigned lit = d.mylit;
            for (int w = 1; w < NW; w++) lit |= s_lit[w][lane];
            const unsigned sure = (d.isdyn && !k.debug_no_vis) ? __float_as_uint(d.rb.w) : 0u;
            lit |= sure & (d.nlights >= 32 ? 0xffffffffu : ((1u << d.nlights) - 1u));
            dyn_finish(k, d, lit, lane);
        }
        __syncthreads();                                       // s_lit and s_next are reused by the next entry
        if (STATS && k.stats && threadIdx.x == 0) {
            const long long dt = clock64() - t_e0;
            t_entries += dt;
            atomicMax(k.stats + STAT_DYN_MAXCYC, (unsigned long long)dt);
            if (dt > 20000) atomicAdd(k.stats + STAT_DYN_SLOW, 1ull);
        }
        ei = ei_next;
    }
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_DYN_RAYS, (unsigned long long)dyn_rays);
        atomicAdd(k.stats + STAT_DYN_ITERS, (unsigned long long)dyn_iters);
        atomicAdd(k.stats + STAT_DYN_ENTRIES, (unsigned long long)dyn_entries);
        if (warp == 0) {
            atomicAdd(k.stats + STAT_DYN_CYCLES, (unsigned long long)t_entries);
            atomicMax(k.stats + STAT_DYN_WARPMAX, (unsigned long long)t_entries);
            atomicMax(k.stats + STAT_DYN_KERNEL, (unsigned long long)(clock64() - t_start));
        }
    }
    // the last CTA out re-arms the queue for the next step (every CTA of view_kernel has signed off by then; the wait
    // makes this grid's completion formally imply that grid's)
    griddep_wait();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(k.dyn_ctrl + 1, 1) == (int)gridDim.x - 1) { k.dyn_ctrl[0] = 0; k.dyn_ctrl[1] = 0; k.dyn_ctrl[2] = 0; k.dyn_ctrl[3] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tick_kernel: view_kernel as a persistent grid — one CTA per resident slot, looping over environments — with the
// second pass (dyn_kernel's work) optionally taken on by the same warps. OPT-IN (option "persist" = 1): built to hide the
// one-shot CTAs' prologue / epilogue stalls, measured slower than view_kernel + dyn_kernel at the benchmark's size
// (DESIGN.md §4 has the counters), kept as a tested variant.
//   * each CTA keeps a POOL of S shared-memory STAGES, one env each. The warp that finishes the last item of the env in
//     a stage restages it: the next env's descriptor (fetched off a global counter, costliest first, while the previous
//     tenant was being worked on, its table prefetched into L2), three bulk (TMA) copies on the stage's mbarrier, the
//     agents loaded and drawn, the stage published — while the other warps are busy with the other stages. No CTA-wide
//     barrier anywhere after the first;
//   * a warp looking for work takes an item of the staged env that was published first and still has items to hand
//     out: an item that takes long keeps its own stage, the others go on being refilled around it;
//   * merged second pass: every warp holds a TICKET for one (queue entry, light group) and looks at that entry's flag
//     between two ray items — a published entry is lit by the `dyn_groups` warps holding its tickets (their finds OR-ed
//     into the entry, the last one sums the lights and writes the pixels); once the rays are done the warps drain what
//     is left. The entries are counted per ENV (dyn_ctrl[3]), so nothing waits on a CTA that is yet to be scheduled.
// Results are those of view_kernel + dyn_kernel, bit for bit (same item code, same per-entry scan).
// ---------------------------------------------------------------------------------------------------------------
enum { PMAXS = 8, META_N = 16, META_G0 = 17, META_L = 19, META_PAR = 20, META_INTS = 24 };
enum { SPIN_LIMIT = 1 << 22 };

struct PCtl {
    uint64_t full_bar[PMAXS];   // per stage: the bulk copies of its env's table have landed
    int ready_seq[PMAXS];       // per stage: 0 while it is being (re)staged; else the publication number of the env staged and
                                // drawn there (older envs first); 0x7fffffff: none will come
    int item_next[PMAXS];       // per stage: items of its env handed out
    int done_cnt[PMAXS];        // per stage: items of its env finished
    int fills[PMAXS];           // per stage: bulk-copy rounds so far (the mbarrier's phase)
    int nseq[PMAXS];            // per stage: nmeta holds the descriptor of its next tenant
    int nmeta[PMAXS][META_INTS];// per stage: that descriptor, fetched while the current tenant is being worked on
    int fetches;                // descriptors fetched by this CTA
    int pubs;                   // envs published by this CTA
    int dead;                   // stages that will not be refilled
    int all_out;                // warps that have left the kernel
    int mrad;                   // bits of the agent model's radius
    int error;                  // a bounded wait ran out (a bug): results are void, nothing hangs
    int pad[3];
};

__device__ __forceinline__ size_t pctl_bytes() { return (sizeof(PCtl) + 15) & ~size_t(15); }

static size_t pstage_bytes(int wcap, int A, int AF, bool stage_rec) {
    size_t b = (size_t)(AF + wcap) * 16 + (size_t)(wcap / VRUN) * 16 + (stage_rec ? (size_t)wcap * 16 : 0) +
               (size_t)A * ST_STRIDE * 4 + (size_t)META_INTS * 4;
    return (b + 15) & ~size_t(15);
}

__device__ __forceinline__ VSmem pstage(unsigned char* base, PCtl* ctl, const KArgs& k, int A, int AF) {
    VSmem m;
    m.seg = reinterpret_cast<float4*>(base);
    m.boxes = m.seg + AF + k.wcap;
    m.rec = reinterpret_cast<int4*>(m.boxes + k.wcap / VRUN);
    m.rec_g = nullptr;
    m.scr = nullptr;
    m.st_in = m.st_out = reinterpret_cast<float*>(m.rec + (k.stage_rec ? k.wcap : 0));
    m.meta = reinterpret_cast<int*>(m.st_out + A * ST_STRIDE);
    m.mrad = &ctl->mrad;
    m.bar = nullptr;
    m.next_item = nullptr;
    return m;
}

__device__ __forceinline__ int ld_volatile_shared(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ int ld_volatile_global(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// An env's DESCRIPTOR: which env comes next and everything static the stage needs to know about it, into `meta`
// ([META_N] = -1: there is none left). Three dependent global reads — why it is fetched one tenancy ahead of its use —
// plus an L2 prefetch of the env's table, so that the bulk copies that will stage it find it on chip.
__device__ __forceinline__ void desc_fetch(const KArgs& k, int* meta, int q, int lane, int AF) {
    int n = -1;
    if (lane == 0) {
        long long e;
        if (q < k.p_stages || !k.sched) e = (long long)blockIdx.x + (long long)q * gridDim.x;     // every CTA's first envs: dealt
        else e = (long long)k.p_stages * gridDim.x + atomicAdd(k.sched, 1);                       // then first come, first served
        if (e < k.s.n_envs) n = k.env_order ? __ldg(k.env_order + e) : (int)e;
        meta[META_N] = n;
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    if (n >= 0) {
        if (lane == 0) {
            const int L = __ldg(k.s.line_widths + n);
            const int64_t g0 = __ldg(k.s.line_starts + n);
            const int b0 = __ldg(k.s.box_starts + n);
            const int W = L - AF, nb = (W + VRUN - 1) / VRUN;
            meta[0] = W; meta[3] = b0;
            meta[META_G0] = (int)(uint32_t)g0; meta[META_G0 + 1] = (int)(uint32_t)((uint64_t)g0 >> 32);
            meta[META_L] = L;
            if (nb > 0 && q >= k.p_stages) {
                bulk_prefetch_l2(k.s.occ_lines + 4 * VRUN * (int64_t)b0, (uint32_t)nb * VRUN * 16u);
                if (k.stage_rec) bulk_prefetch_l2(k.s.occ_rec + 4 * VRUN * (int64_t)b0, (uint32_t)nb * VRUN * 16u);
                bulk_prefetch_l2(k.s.occ_boxes + 4 * (int64_t)b0, (uint32_t)nb * 16u);
            }
        } else if (lane == 1) {
            meta[1] = __ldg(k.s.light_widths + n); meta[2] = __ldg(k.s.light_starts + n);
        } else if (lane == 2) {
            meta[4] = __float_as_int(__ldg(k.s.occ_meta + 2 * n)); meta[5] = __float_as_int(__ldg(k.s.occ_meta + 2 * n + 1));
            meta[6] = 0; meta[7] = 0;
        } else if (lane == 3) {
            meta[10] = 0; meta[11] = 0;                                  // no grid: every lookup falls outside
            if (k.s.vis) {
                const float4 vm = __ldg(reinterpret_cast<const float4*>(k.s.vis_meta) + n);
                const int64_t vs = __ldg(k.s.vis_starts + n);
                meta[8] = __float_as_int(vm.x); meta[9] = __float_as_int(vm.y); meta[10] = (int)vm.z; meta[11] = (int)vm.w;
                meta[12] = (int)(uint32_t)vs; meta[13] = (int)(uint32_t)((uint64_t)vs >> 32);
            }
        }
    }
    __syncwarp();
}

// The env described in the stage's meta block: its table on its way into the stage by three bulk copies (static scenery
// only: may run ahead of griddep_wait).
__device__ __forceinline__ void stage_load(const KArgs& k, PCtl* ctl, const VSmem& m, int s, int lane, int AF) {
    if (lane == 0) {
        const int W = m.meta[0], b0 = m.meta[3], nb = (W + VRUN - 1) / VRUN;
        m.meta[META_PAR] = ctl->fills[s] & 1;
        if (nb > 0) {
            ctl->fills[s]++;
            // the stage's previous tenant was read through the generic proxy: order those reads (synchronised with by the
            // done counter) before the async-proxy writes of the copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&ctl->full_bar[s], (uint32_t)nb * (VRUN * (k.stage_rec ? 32u : 16u) + 16u));
            bulk_g2s(m.seg + AF, k.s.occ_lines + 4 * VRUN * (int64_t)b0, (uint32_t)nb * VRUN * 16u, &ctl->full_bar[s]);
            if (k.stage_rec) bulk_g2s(m.rec, k.s.occ_rec + 4 * VRUN * (int64_t)b0, (uint32_t)nb * VRUN * 16u, &ctl->full_bar[s]);
            bulk_g2s(m.boxes, k.s.occ_boxes + 4 * (int64_t)b0, (uint32_t)nb * 16u, &ctl->full_bar[s]);
        }
    }
    __syncwarp();
}

// Second half (needs the kernels ahead in the stream: call after griddep_wait): the agents' state and their model lines
// at their current poses (draw_kernel, kernels.cu:297-318), the IMU head; then wait for the table and publish the stage.
__device__ __forceinline__ void stage_publish(const KArgs& k, PCtl* ctl, const VSmem& m, int s, int n, int lane, int A) {
    const int F = k.s.n_model;
    for (int a = lane; a < A; a += 32) {
        const int64_t i = (int64_t)n * A + a;
        const float2 pos = reinterpret_cast<const float2*>(k.a.positions)[i];
        const float2 vel = reinterpret_cast<const float2*>(k.a.velocity)[i];
        const float ang = k.a.angles[i];
        const float av = k.a.angvelocity[i];
        float* st = m.st_out + a * ST_STRIDE;
        st[ST_ANG] = ang; st[ST_PX] = pos.x; st[ST_PY] = pos.y; st[ST_AV] = av; st[ST_VX] = vel.x; st[ST_VY] = vel.y;
        sincos_deg(ang, st[ST_SN], st[ST_CS]);                  // once per agent (kernels.cu:304-306, 335-337 redo it per thread)
        if (k.has_obs && k.obs.imu) {                             // IMU (modules.py:263-270)
            const float rad = __fmul_rn(0.017453292519943295f, ang);
            const float c = cosf(rad), sn = sinf(rad);
            float* q3 = k.obs.imu + 3 * i;
            q3[0] = __fmul_rn(av, k.inv_ang);
            q3[1] = __fmul_rn(__fadd_rn(__fmul_rn(c, vel.x), __fmul_rn(sn, vel.y)), k.inv_speed);
            q3[2] = __fmul_rn(__fadd_rn(__fmul_rn(-sn, vel.x), __fmul_rn(c, vel.y)), k.inv_speed);
        }
    }
    __syncwarp();
    const int64_t g0 = (int64_t)(((uint64_t)(uint32_t)m.meta[META_G0 + 1] << 32) | (uint32_t)m.meta[META_G0]);
    for (int t = lane; t < 2 * F * A; t += 32) {                // t = 2 * (agent's model line) + endpoint, over all agents
        const int a = t / (2 * F), tt = t - a * 2 * F;
        const float* st = m.st_out + a * ST_STRIDE;
        const float sn = st[ST_SN], cs = st[ST_CS];
        const float2 mp = __ldg(reinterpret_cast<const float2*>(k.s.model) + tt);
        const float2 pt = make_float2(fadd(st[ST_PX], cross2(cs, mp.x, sn, mp.y)), fadd(st[ST_PY], dot2(sn, mp.x, cs, mp.y)));
        reinterpret_cast<float2*>(m.seg)[t] = pt;
        reinterpret_cast<float2*>(k.s.lines)[2 * g0 + t] = pt;
    }
    if (m.meta[0] > 0) mbar_wait(&ctl->full_bar[s], (uint32_t)m.meta[META_PAR]);
    __syncwarp();
    if (lane == 0) {
        const int pub = atomicAdd(&ctl->pubs, 1) + 1;
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&ctl->ready_seq[s]) = pub;
    }
}

__device__ __forceinline__ void stage_dead(PCtl* ctl, int s, int lane) {
    if (lane == 0) {
        atomicAdd(&ctl->dead, 1);
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&ctl->ready_seq[s]) = 0x7fffffff;
    }
}

// One ticket of the merged second pass: entry T / groups, lights i = T % groups (mod groups). Returns after the entry's
// share is done; the last of the entry's `groups` warps to finish sums the lights and writes the pixels.
template <bool STATS>
__device__ __noinline__ void dyn_ticket_run(const KArgs& k, int T, int lane) {
    const int G = k.dyn_groups, PS = k.dyn_window;
    const int ei = T / G, part = T - ei * G;
    unsigned char* e = k.dyn_entries + (size_t)ei * (DYN_HDR + 32 * PS);
    DynEntry d;
    unsigned iters = 0;
    dyn_scan<STATS>(k, e, part, G, lane, d, iters);
    float4* rec = reinterpret_cast<float4*>(e + DYN_HDR) + 2 * lane;
    int* parts_done = reinterpret_cast<int*>(e + 40);
    if (G > 1) {
        // my finds join the entry: OR-ed into the word that already holds the lights the visibility grid vouches for
        if (d.isdyn) {
            if (d.nlights <= 32) { if (d.mylit) atomicOr(reinterpret_cast<unsigned*>(&rec[1].w), d.mylit); }
            else if (part == 0) __stcg(&rec[1].z, d.intensity);
        }
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence(); last = atomicAdd(parts_done, 1) == G - 1; __threadfence(); }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return;
        if (d.isdyn) {
            d.rb.w = __ldcg(&rec[1].w);
            if (d.nlights > 32) d.intensity = __ldcg(&rec[1].z);
        }
    }
    unsigned lit = d.mylit;
    if (d.isdyn) {
        const unsigned bits = __float_as_uint(d.rb.w);                       // the grid's lights (+ the other parts' finds)
        lit |= bits & (d.nlights >= 32 ? 0xffffffffu : ((1u << d.nlights) - 1u));
    }
    dyn_finish(k, d, lit, lane);
    __syncwarp();
    if (lane == 0) {
        *parts_done = 0;
        *reinterpret_cast<volatile int*>(e + DYN_FLAG) = 0;                  // re-armed for the next step
    }
}

template <int NCH, bool MERGE, bool STATS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) tick_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int S = k.p_stages, IPE = k.p_items;
    PCtl* ctl = reinterpret_cast<PCtl*>(smem_raw);
    unsigned char* stages = smem_raw + pctl_bytes();
    float4* scr = reinterpret_cast<float4*>(stages + (size_t)S * k.p_stage_bytes) + warp * (32 * NCH < 64 ? 64 : 32 * NCH);
    if (tid == 0) {
        for (int s = 0; s < PMAXS; s++) { ctl->ready_seq[s] = 0; ctl->item_next[s] = 0; ctl->done_cnt[s] = 0; ctl->fills[s] = 0; ctl->nseq[s] = 0; }
        ctl->fetches = S; ctl->pubs = 0; ctl->dead = 0; ctl->all_out = 0; ctl->error = 0;
        float r = 0.f;
        for (int t = 0; t < 2 * k.s.n_model; t++) {                          // the model's radius: lets a warp skip agents it cannot see
            const float2 mp = __ldg(reinterpret_cast<const float2*>(k.s.model) + t);
            r = fmaxf(r, sqrtf(mp.x * mp.x + mp.y * mp.y));
        }
        ctl->mrad = __float_as_int(r);
        for (int s = 0; s < S; s++) mbar_init(&ctl->full_bar[s], 1);
    }
    __syncthreads();
    // the first S envs: their tables start moving before the kernels ahead in the stream are known to be done
    int n0 = -1;
    VSmem m0;
    if (warp < S) {
        m0 = pstage(stages + (size_t)warp * k.p_stage_bytes, ctl, k, A, AF);
        desc_fetch(k, m0.meta, warp, lane, AF);
        n0 = m0.meta[META_N];
        if (n0 >= 0) stage_load(k, ctl, m0, warp, lane, AF);
    }
    griddep_wait();
    griddep_launch();
    if (warp < S) {
        if (n0 >= 0) stage_publish(k, ctl, m0, warp, n0, lane, A);
        else stage_dead(ctl, warp, lane);
        // and the descriptor of the stage's next tenant
        int f = 0;
        if (lane == 0) f = atomicAdd(&ctl->fetches, 1);
        desc_fetch(k, ctl->nmeta[warp], __shfl_sync(0xffffffffu, f, 0), lane, AF);
        if (lane == 0) { __threadfence_block(); *reinterpret_cast<volatile int*>(&ctl->nseq[warp]) = 1; }
    }
    const bool merge = MERGE && k.dyn_entries != nullptr;
    const int esize = DYN_HDR + 32 * k.dyn_window;
    int T = -1, fl = 0;                                             // my ticket of the second pass; its entry's flag at the last look
    auto take_ticket = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(k.dyn_ctrl + 2, 1);
        T = __shfl_sync(0xffffffffu, t, 0);
    };
    auto peek = [&]() -> int {                                      // (every lane reads the same word: one transaction, no shuffle)
        const int ei = T / k.dyn_groups;
        return ei < k.dyn_cap ? ld_volatile_global(reinterpret_cast<const int*>(k.dyn_entries + (size_t)ei * esize + DYN_FLAG)) : 0;
    };
    if (merge) { take_ticket(); fl = peek(); }
    bool rays = true;
    unsigned spins = 0;
    long long c_items = 0, c_wait = 0, c_dyn = 0, c_drain = 0, c_prep = 0, c_fetch = 0, n_dyn = 0, n_items = 0, n_spun = 0;
    const long long c_start = STATS ? clock64() : 0;
    for (;;) {
        if (merge && __any_sync(0xffffffffu, fl != 0)) {
            const long long c0 = STATS ? clock64() : 0;
            dyn_ticket_run<STATS>(k, T, lane);
            take_ticket();
            fl = peek();
            if (STATS) { c_dyn += clock64() - c0; n_dyn++; }
            continue;
        }
        if (rays) {
            // an item of the staged env that was published first and still has items to hand out. (The stages are a pool,
            // not a ring: an item that takes long keeps ITS stage; the others go on being refilled around it.)
            const long long c0 = STATS ? clock64() : 0;
            int s = -1, alive = 0, oldest = 0x7fffffff;
            for (int i = 0; i < S; i++) {
                const int v = ld_volatile_shared(&ctl->ready_seq[i]);
                if (v == 0x7fffffff) continue;
                alive++;
                if (v > 0 && v < oldest && ld_volatile_shared(&ctl->item_next[i]) < IPE) { oldest = v; s = i; }
            }
            if (s < 0) {
                if (alive == 0 || ld_volatile_shared(&ctl->error)) {        // nothing is staged and nothing will be: the rays are done
                    rays = false;
                    if (!merge) break;
                    continue;
                }
                __nanosleep(200);
                if (STATS) { c_wait += clock64() - c0; n_spun++; }
                if (++spins > SPIN_LIMIT) { ctl->error = 1; if (k.sched) k.sched[2] = 1; }
                continue;
            }
            int it = 0;
            if (lane == 0) it = atomicAdd(&ctl->item_next[s], 1);
            it = __reduce_max_sync(0xffffffffu, it);                // (lands in a uniform register: so do the item's address parts)
            if (it >= IPE) continue;                                // another warp was quicker
            // (the stage may have been finished and taken down for restaging since I looked: my item is then one of its
            // next tenant's, or void)
            int v;
            while ((v = ld_volatile_shared(&ctl->ready_seq[s])) == 0) {
                __nanosleep(200);
                if (++spins > SPIN_LIMIT) { ctl->error = 1; if (k.sched) k.sched[2] = 1; v = 0x7fffffff; break; }
            }
            if (v == 0x7fffffff) continue;
            __threadfence_block();
            const long long c1 = STATS ? clock64() : 0;
            VSmem m = pstage(stages + (size_t)s * k.p_stage_bytes, ctl, k, A, AF);
            const int fl_next = merge ? peek() : 0;                 // asked for now, looked at after the item
            {
                const int n = __reduce_max_sync(0xffffffffu, m.meta[META_N]);
                const int64_t g0 = (int64_t)(((uint64_t)(uint32_t)m.meta[META_G0 + 1] << 32) | (uint32_t)m.meta[META_G0]);
                const int L = m.meta[META_L], W = m.meta[0], nb = (W + VRUN - 1) / VRUN;
                m.rec_g = reinterpret_cast<const int4*>(k.s.occ_rec) + VRUN * (int64_t)m.meta[3];
                const int RB = k.ray_blocks, rbs = k.rb_shift;
                const int a = rbs >= 0 ? it >> rbs : it / RB;
                view_agent<NCH, STATS>(k, m, n, g0, L, W, nb, a, it - a * RB, scr, lane);
            }
            fl = fl_next;
            __syncwarp();
            const long long c2 = STATS ? clock64() : 0;
            if (STATS) { c_items += c2 - c1; n_items++; }
            int last = 0;
            if (lane == 0) { __threadfence_block(); last = atomicAdd(&ctl->done_cnt[s], 1) == IPE - 1; }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                // the env in stage s is finished (nothing more will be queued for it: the second pass counts envs, so
                // that it depends on no CTA that is yet to be scheduled): take the stage down and stage the next env there
                if (lane == 0) {
                    *reinterpret_cast<volatile int*>(&ctl->ready_seq[s]) = 0;
                    __threadfence_block();
                    ctl->done_cnt[s] = 0;
                    *reinterpret_cast<volatile int*>(&ctl->item_next[s]) = 0;
                    if (k.dyn_entries) { __threadfence(); red_release_add(k.dyn_ctrl + 3, 1); }
                }
                // its descriptor was fetched while this env was being worked on
                while (ld_volatile_shared(&ctl->nseq[s]) == 0) {
                    __nanosleep(100);
                    if (++spins > SPIN_LIMIT) { ctl->error = 1; if (k.sched) k.sched[2] = 1; break; }
                }
                __threadfence_block();
                if (lane < META_INTS) m.meta[lane] = ctl->nmeta[s][lane];
                __syncwarp();
                const int n = m.meta[META_N];
                if (n >= 0) {
                    if (lane == 0) ctl->nseq[s] = 0;
                    stage_load(k, ctl, m, s, lane, AF);
                    stage_publish(k, ctl, m, s, n, lane, A);
                    const long long c3 = STATS ? clock64() : 0;
                    int f = 0;
                    if (lane == 0) f = atomicAdd(&ctl->fetches, 1);
                    desc_fetch(k, ctl->nmeta[s], __shfl_sync(0xffffffffu, f, 0), lane, AF);
                    if (STATS) { c_fetch += clock64() - c3; c_prep -= clock64() - c3; }
                    if (lane == 0) { __threadfence_block(); *reinterpret_cast<volatile int*>(&ctl->nseq[s]) = 1; }
                } else {
                    stage_dead(ctl, s, lane);
                }
                if (STATS) c_prep += clock64() - c2;
            }
        } else {
            // drain: my entry will be published, or every env is done and it never will be
            const long long c0 = STATS ? clock64() : 0;
            fl = peek();
            if (__any_sync(0xffffffffu, fl != 0)) continue;
            if (ld_volatile_global(k.dyn_ctrl + 3) >= k.s.n_envs) {
                __threadfence();
                fl = peek();
                if (!__any_sync(0xffffffffu, fl != 0)) break;
                continue;
            }
            __nanosleep(200);
            if (STATS) c_drain += clock64() - c0;
            if (++spins > 16u * SPIN_LIMIT) { if (k.sched) k.sched[2] = 1; break; }         // (seconds: a bug, not a wait)
        }
    }
    if (STATS && k.stats && lane == 0) {
        const long long total = clock64() - c_start;
        atomicAdd(k.stats + STAT_T_ITEMS, (unsigned long long)c_items); atomicAdd(k.stats + STAT_T_WAIT, (unsigned long long)c_wait);
        atomicAdd(k.stats + STAT_T_DYN, (unsigned long long)c_dyn); atomicAdd(k.stats + STAT_T_DRAIN, (unsigned long long)c_drain);
        atomicAdd(k.stats + STAT_T_PREP, (unsigned long long)c_prep); atomicAdd(k.stats + STAT_T_TOTAL, (unsigned long long)total);
        atomicAdd(k.stats + STAT_N_DYN, (unsigned long long)n_dyn); atomicAdd(k.stats + STAT_N_ITEMS, (unsigned long long)n_items);
        atomicMax(k.stats + STAT_T_MAXWARP, (unsigned long long)total); atomicAdd(k.stats + STAT_N_SPUN, (unsigned long long)n_spun);
        atomicAdd(k.stats + STAT_T_FETCH, (unsigned long long)c_fetch);
    }
    // the last warp of the last CTA out re-arms the counters for the next launch
    __syncwarp();
    if (lane == 0 && atomicAdd(&ctl->all_out, 1) == nwarps - 1 && k.sched) {
        __threadfence();
        if (atomicAdd(k.sched + 1, 1) == (int)gridDim.x - 1) {
            k.sched[0] = 0; k.sched[1] = 0;
            if (merge) { k.dyn_ctrl[0] = 0; k.dyn_ctrl[1] = 0; k.dyn_ctrl[2] = 0; k.dyn_ctrl[3] = 0; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// table_kernel: the spatial table (msb_build_table). One CTA per env; two bitonic sorts of 64-bit keys in shared memory.
// Sort-tile-recursive packing: the static segments are ranked by the x of their midpoints, cut into ~sqrt(nb) strips of
// whole runs, each strip sorted by y; consecutive 16 form a run. Ties break by line index (a stable sort).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sortable(float f) {          // bit pattern that orders like the float
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

__device__ __forceinline__ void bitonic_sort(uint64_t* keys, int n2) {
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int o = i ^ j;
                if (o > i) {
                    const uint64_t a = keys[i], b = keys[o];
                    if ((a > b) == ((i & kk) == 0)) { keys[i] = b; keys[o] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) table_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ float red[8][5];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AF = k.s.n_agents * k.s.n_model;
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = (int64_t)__ldg(k.s.line_starts + n) + AF;
    const int W = L > AF ? L - AF : 0;
    const int nb = (W + VRUN - 1) / VRUN;
    const int64_t b0 = __ldg(k.s.box_starts + n);
    const float4* lines = reinterpret_cast<const float4*>(k.s.lines) + g0;
    float4* occ = const_cast<float4*>(reinterpret_cast<const float4*>(k.s.occ_lines)) + VRUN * b0;
    int4* rec = const_cast<int4*>(reinterpret_cast<const int4*>(k.s.occ_rec)) + VRUN * b0;
    float4* boxes = const_cast<float4*>(reinterpret_cast<const float4*>(k.s.occ_boxes)) + b0;
    int n2 = 1;
    while (n2 < W) n2 <<= 1;
    // 1. by (mid x, line)
    float vmax = 0.f, lox = CUDART_INF_F, loy = CUDART_INF_F, hix = -CUDART_INF_F, hiy = -CUDART_INF_F;
    for (int i = tid; i < n2; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < W) {
            const float4 s4 = __ldg(lines + i);
            key = ((uint64_t)sortable((s4.x + s4.z) * .5f) << 32) | (uint32_t)i;
            vmax = fmaxf(vmax, fmaxf(fabsf(s4.z - s4.x), fabsf(s4.w - s4.y)));
            lox = fminf(lox, fminf(s4.x, s4.z)); hix = fmaxf(hix, fmaxf(s4.x, s4.z));
            loy = fminf(loy, fminf(s4.y, s4.w)); hiy = fmaxf(hiy, fmaxf(s4.y, s4.w));
        }
        keys[i] = key;
    }
    __syncthreads();
    bitonic_sort(keys, n2);
    // 2. by (strip of the x ranking, mid y, line): 10 + 32 + 14 bits (msb_build_table refuses more than 16383 lines)
    const int strips = nb > 0 ? (int)ceil(sqrt((double)nb)) : 1;
    const int per_strip = ((nb + strips - 1) / strips) * VRUN;
    for (int r = tid; r < W; r += blockDim.x) {
        const uint32_t i = (uint32_t)keys[r];
        const float4 s4 = __ldg(lines + i);
        const uint32_t strip = (uint32_t)(r / (per_strip > 0 ? per_strip : 1));
        keys[r] = ((uint64_t)strip << 46) | ((uint64_t)sortable((s4.y + s4.w) * .5f) << 14) | i;
    }
    __syncthreads();
    bitonic_sort(keys, n2);
    // 3. rows (padding: a far, zero-length segment with line index -1), run boxes, per-env summary
    for (int p = tid; p < nb * VRUN; p += blockDim.x) {
        float4 s4 = make_float4(1e30f, 1e30f, 1e30f, 1e30f);
        int4 rc = make_int4(0, 0, 0, -1);
        if (p < W) {
            const int i = (int)(keys[p] & 0x3fffu);
            s4 = __ldg(lines + i);
            const int64_t ts = __ldg(k.s.tex_starts + g0 + i);
            rc = make_int4((int)(uint32_t)ts, (int)(uint32_t)((uint64_t)ts >> 32), __ldg(k.s.tex_widths + g0 + i), AF + i);
        }
        occ[p] = s4;
        rec[p] = rc;
    }
    for (int b = tid; b < nb; b += blockDim.x) {
        float x0 = CUDART_INF_F, y0 = CUDART_INF_F, x1 = -CUDART_INF_F, y1 = -CUDART_INF_F;
        for (int p = b * VRUN; p < (b + 1) * VRUN && p < W; p++) {
            const float4 s4 = __ldg(lines + (int)(keys[p] & 0x3fffu));
            x0 = fminf(x0, fminf(s4.x, s4.z)); x1 = fmaxf(x1, fmaxf(s4.x, s4.z));
            y0 = fminf(y0, fminf(s4.y, s4.w)); y1 = fmaxf(y1, fmaxf(s4.y, s4.w));
        }
        boxes[b] = make_float4(x0, y0, x1, y1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    if (lane == 0) { red[warp][0] = vmax; red[warp][1] = lox; red[warp][2] = loy; red[warp][3] = hix; red[warp][4] = hiy; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            vmax = fmaxf(vmax, red[w][0]); lox = fminf(lox, red[w][1]); loy = fminf(loy, red[w][2]);
            hix = fmaxf(hix, red[w][3]); hiy = fmaxf(hiy, red[w][4]);
        }
        float* meta = const_cast<float*>(k.s.occ_meta) + 2 * n;
        meta[0] = vmax;
        meta[1] = W > 0 ? fmaxf(hix - lox, hiy - loy) : 0.f;
        const_cast<int32_t*>(k.s.occ_starts)[n] = (int32_t)(VRUN * b0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// vis_kernel: the light-visibility grid (msb_scenery::vis). One CTA per env, one warp per cell, lane = light.
// A light is vouched for when NO static segment comes within `thr` of the segment light -> cell centre, where
//   thr = mg + half the cell's diagonal + slack,   mg = delta * (|U| + vmax) + 0.01,   delta = 4e-4 * vmax * (diam + |U|)
// is dyn_kernel's own shadow-cull margin (DESIGN.md "shadow cull": the reference's intersect() can only report a
// crossing if the two segments truly come within delta * (|U| + |V|) of each other) evaluated for the farthest point of
// the cell. Every segment light -> point-of-the-cell lies within half a diagonal of light -> centre, so for no point of
// the cell can the reference's shadow test fire. The slack covers float rounding here and in the cell lookup.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pt_seg_d2(float x, float y, float px, float py, float dx, float dy) {
    const float wx = x - px, wy = y - py;
    const float t = fminf(fmaxf((wx * dx + wy * dy) / fmaxf(dx * dx + dy * dy, 1e-30f), 0.f), 1.f);
    const float ex = wx - t * dx, ey = wy - t * dy;
    return ex * ex + ey * ey;
}

__global__ void __launch_bounds__(256) vis_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int AF = k.s.n_agents * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float4* seg = reinterpret_cast<float4*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(seg + k.wcap);
    const int W = __ldg(k.s.line_widths + n) - AF;
    const int nb = (W + VRUN - 1) / VRUN;
    if (tid == 0) {
        mbar_init(bar, 1);
        if (nb > 0) {
            mbar_expect_tx(bar, (uint32_t)nb * VRUN * 16u);
            bulk_g2s(seg, k.s.occ_lines + 4 * VRUN * (int64_t)__ldg(k.s.box_starts + n), (uint32_t)nb * VRUN * 16u, bar);
        }
    }
    __syncthreads();
    if (nb > 0) mbar_wait(bar, 0);
    const float4 vm = __ldg(reinterpret_cast<const float4*>(k.s.vis_meta) + n);
    const int gx = (int)vm.z, gy = (int)vm.w;
    uint32_t* out = k.s.vis + __ldg(k.s.vis_starts + n);
    const int I = __ldg(k.s.light_widths + n);
    const float* lt = k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n);
    const float vmax = __ldg(k.s.occ_meta + 2 * n), diam = __ldg(k.s.occ_meta + 2 * n + 1);
    const float4* boxes = reinterpret_cast<const float4*>(k.s.occ_boxes) + __ldg(k.s.box_starts + n);
    const bool real = lane < I;
    const float Ix = real ? __ldg(lt + 3 * lane) : 0.f, Iy = real ? __ldg(lt + 3 * lane + 1) : 0.f;
    for (int cell = warp; cell < gx * gy; cell += nwarps) {
        const int iy = cell / gx, ix = cell - iy * gx;
        const float Cx = vm.x + ((float)ix + 0.5f) * VIS_CELL, Cy = vm.y + ((float)iy + 0.5f) * VIS_CELL;
        const float Ux = Cx - Ix, Uy = Cy - Iy;
        const float ulen = fmaxf(fabsf(Ux), fabsf(Uy)) + 0.5f * VIS_CELL;          // as dyn_kernel measures |U|, at the cell's far corner
        const float delta = 4e-4f * vmax * (diam + ulen);
        const float thr = delta * (ulen + vmax) + 0.01f + 0.70711f * VIS_CELL + 0.01f;
        const float thr2 = thr * thr;
        const float bx0 = fminf(Ix, Cx) - thr, bx1 = fmaxf(Ix, Cx) + thr, by0 = fminf(Iy, Cy) - thr, by1 = fmaxf(Iy, Cy) + thr;
        bool sure = real;
        for (int b = 0; b < nb; b++) {
            // the run's box against the box around light -> centre: a run no light's query comes near is skipped whole
            const float4 bx = __ldg(boxes + b);                                     // (uniform address: a broadcast)
            const bool near = sure && !(bx.z < bx0 || bx.x > bx1 || bx.w < by0 || bx.y > by1);
            if (!__any_sync(0xffffffffu, near)) continue;
            for (int l = VRUN * b; l < VRUN * (b + 1) && l < W; l++) {
                const float4 s4 = seg[l];                                           // (uniform address: a broadcast)
                // nowhere near the box around light -> centre: the common case
                const bool far = fmaxf(s4.x, s4.z) < bx0 || fminf(s4.x, s4.z) > bx1 || fmaxf(s4.y, s4.w) < by0 || fminf(s4.y, s4.w) > by1;
                if (sure && !far) {
                    const float Vx = s4.z - s4.x, Vy = s4.w - s4.y;
                    // do they cross? (orientation signs; a near-miss shows up in the end-point distances below)
                    const float o1 = Ux * (s4.y - Iy) - Uy * (s4.x - Ix), o2 = Ux * (s4.w - Iy) - Uy * (s4.z - Ix);
                    const float o3 = Vx * (Iy - s4.y) - Vy * (Ix - s4.x), o4 = Vx * (Cy - s4.y) - Vy * (Cx - s4.x);
                    const bool cross = (o1 * o2 <= 0.f) && (o3 * o4 <= 0.f);
                    const float d2 = fminf(fminf(pt_seg_d2(s4.x, s4.y, Ix, Iy, Ux, Uy), pt_seg_d2(s4.z, s4.w, Ix, Iy, Ux, Uy)),
                                           fminf(pt_seg_d2(Ix, Iy, s4.x, s4.y, Vx, Vy), pt_seg_d2(Cx, Cy, s4.x, s4.y, Vx, Vy)));
                    if (cross || !(d2 > thr2)) sure = false;                        // (NaN: not sure)
                }
            }
            if (!__any_sync(0xffffffffu, sure)) break;
        }
        const unsigned word = __ballot_sync(0xffffffffu, sure);
        if (lane == 0) out[cell] = word;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bake (kernels.cu:270-293): one CTA per env, one warp per line, one lane per texel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bake_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float4* seg = reinterpret_cast<float4*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(seg + k.seg_cap);
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    if (tid == 0) {
        mbar_init(bar, 1);
        if (L > 0) {
            mbar_expect_tx(bar, (uint32_t)L * 16u);
            bulk_g2s(seg, k.s.lines + 4 * g0, (uint32_t)L * 16u, bar);
        }
    }
    __syncthreads();
    if (L > 0) mbar_wait(bar, 0);
    const int I = k.s.light_widths[n];
    const float* lt = k.s.lights + 3 * (int64_t)k.s.light_starts[n];
    for (int l = warp; l < L; l += nwarps) {
        const int w = __ldg(k.s.tex_widths + g0 + l);
        const int64_t ts = __ldg(k.s.tex_starts + g0 + l);
        const float4 s4 = seg[l];
        const float rw = rcp((float)w);
        for (int t = lane; t < w; t += 32) {
            const float loc = fmul(fadd((float)(unsigned)t, 0.5f), rw);                     // :278
            const float om = fsub(1.f, loc);
            const float Cx = ffma(s4.x, om, fmul(loc, s4.z)), Cy = ffma(s4.y, om, fmul(loc, s4.w));   // :279
            k.s.baked[ts + t] = light_intensity_thread(seg, AF, L, lt, I, Cx, Cy);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bake over the spatial table: the same light_intensity() per texel, but a light's shadow test only opens the runs whose
// (conservatively grown) box the segment light -> texel passes through — the slab test and margin of dyn_scan — instead
// of every static line of the env. Which segments get tested varies; occluded / unoccluded per (texel, light) does not,
// and the lights are summed in light order: bit-identical to bake_kernel (and to the reference's baking_kernel).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bake_table_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    float4* seg = reinterpret_cast<float4*>(smem_raw);                   // [wcap] sorted static rows
    float4* boxes = seg + k.wcap;                                        // [wcap / 16]
    uint64_t* bar = reinterpret_cast<uint64_t*>(boxes + k.wcap / VRUN);
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    const int W = L > AF ? L - AF : 0, nb = (W + VRUN - 1) / VRUN;
    if (tid == 0) {
        mbar_init(bar, 1);
        if (nb > 0) {
            const int64_t b0 = __ldg(k.s.box_starts + n);
            mbar_expect_tx(bar, (uint32_t)nb * (VRUN * 16u + 16u));
            bulk_g2s(seg, k.s.occ_lines + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u, bar);
            bulk_g2s(boxes, k.s.occ_boxes + 4 * b0, (uint32_t)nb * 16u, bar);
        }
    }
    __syncthreads();
    if (nb > 0) mbar_wait(bar, 0);
    const int I = __ldg(k.s.light_widths + n);
    const float* lt = k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n);
    const float vmax = __ldg(k.s.occ_meta + 2 * n), diam = __ldg(k.s.occ_meta + 2 * n + 1);
    const float4* lines = reinterpret_cast<const float4*>(k.s.lines) + g0;
    for (int l = warp; l < L; l += nwarps) {
        const int w = __ldg(k.s.tex_widths + g0 + l);
        const int64_t ts = __ldg(k.s.tex_starts + g0 + l);
        const float4 s4 = __ldg(lines + l);
        const float rw = rcp((float)w);
        for (int t = lane; t < w; t += 32) {
            const float loc = fmul(fadd((float)(unsigned)t, 0.5f), rw);                     // kernels.cu:278
            const float om = fsub(1.f, loc);
            const float Cx = ffma(s4.x, om, fmul(loc, s4.z)), Cy = ffma(s4.y, om, fmul(loc, s4.w));   // :279
            float acc = 0.1f;                                                               // AMBIENT
            for (int i = 0; i < I; i++) {
                const float Ix = __ldg(lt + 3 * i), Iy = __ldg(lt + 3 * i + 1), Ii = __ldg(lt + 3 * i + 2);
                const float Ux = fsub(Cx, Ix), Uy = fsub(Cy, Iy);
                // conservative query, as dyn_scan's (DESIGN.md "shadow cull")
                const float ulen = fmaxf(fabsf(Ux), fabsf(Uy));
                const float delta = 4e-4f * vmax * (diam + ulen);
                const float mg = delta * (ulen + vmax) + 0.01f;
                const float irx = 1.f / Ux, iry = 1.f / Uy;                                 // +-inf when axis-parallel
                bool occluded = false;
                for (int b = 0; b < nb && !occluded; b++) {
                    const float4 bx = boxes[b];
                    const float tx0 = (bx.x - mg - Ix) * irx, tx1 = (bx.z + mg - Ix) * irx;
                    const float ty0 = (bx.y - mg - Iy) * iry, ty1 = (bx.w + mg - Iy) * iry;
                    const float tlo = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), 0.f);
                    const float thi = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), 1.f);
                    if (!(tlo > thi)) {                                                     // (NaN: visit)
                        const float4* run = seg + VRUN * b;
#pragma unroll 4
                        for (int r = 0; r < VRUN; r++) {
                            if (occludes(intersect(Ix, Iy, Ux, Uy, run[r]))) { occluded = true; break; }
                        }
                    }
                }
                if (!occluded) {
                    const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
                    acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);
                }
            }
            k.s.baked[ts + t] = fminf(acc, 1.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Environment rules on the device: the game logic the reference's demo envs run as chains of PyTorch
// ops around physics()/render(), each as one small launch with no host round trip.
//   ledger_mark_kernel   Explorer's reward bookkeeping (demo/envs/explorer.py:34-58): which texel every ray landed on, one
//                        bit per texel of the whole scene; the count of bits newly set per env IS the step's potential
//                        gain — instead of a scatter_add over every texel of every env (28 M elements) each step
//   ledger_clear_kernel  envs that reset forget what they have seen (explorer.py:73-77)
//   shoot_kernel         Deathmatch's crosshair rule (demo/envs/deathmatch.py:54-72, 75-80) + the health / damage updates
//   respawn_kernel       RandomSpawns (modules.py:312-326) with the draw on the device (a counter-based hash: no
//                        nonzero(), i.e. no device-to-host sync in the middle of every step)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ledger_mark_kernel(const int32_t* __restrict__ indices, const float* __restrict__ locations,
                                                          const int32_t* __restrict__ line_starts, const int32_t* __restrict__ tex_widths,
                                                          const int64_t* __restrict__ tex_starts, uint32_t* seen, int32_t* potential,
                                                          int32_t* gained, int64_t n_rays, int32_t rays_per_env) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool fresh = false;
    int n = -1;
    if (i < n_rays) {
        n = (int)(i / rays_per_env);
        const int idx = __ldg(indices + i);
        if (idx >= 0) {
            const int64_t g = (int64_t)__ldg(line_starts + n) + idx;
            const float w = (float)__ldg(tex_widths + g);
            // explorer.py:40: min(floor(width * location), width - 1)
            const int64_t texel = __ldg(tex_starts + g) + (int64_t)fminf(floorf(__fmul_rn(w, __ldg(locations + i))), w - 1.f);
            const uint32_t bit = 1u << (texel & 31);
            fresh = !(atomicOr(seen + (texel >> 5), bit) & bit);
        }
    }
    // one add per env and warp (a warp's rays belong to one env, or a few at env boundaries)
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, fresh);
    while (todo) {
        const int env = __shfl_sync(0xffffffffu, n, __ffs(todo) - 1);
        const unsigned same = __ballot_sync(0xffffffffu, fresh && n == env);
        if (lane == __ffs(same) - 1) { atomicAdd(potential + env, __popc(same)); atomicAdd(gained + env, __popc(same)); }
        todo &= ~same;
    }
}

// one CTA per env: if the env resets, its texels' bits — a range of the scene-wide bit array, not word-aligned — are cleared
__global__ void __launch_bounds__(128) ledger_clear_kernel(const uint8_t* __restrict__ reset, const int32_t* __restrict__ line_starts,
                                                           const int32_t* __restrict__ line_widths, const int32_t* __restrict__ tex_widths,
                                                           const int64_t* __restrict__ tex_starts, uint32_t* seen, int32_t* potential) {
    const int n = blockIdx.x;
    if (!reset[n]) return;
    const int64_t g0 = line_starts[n], g1 = g0 + line_widths[n];
    if (g1 == g0) return;
    const int64_t t0 = tex_starts[g0], t1 = tex_starts[g1 - 1] + tex_widths[g1 - 1];
    const int64_t w0 = t0 >> 5, w1 = (t1 + 31) >> 5;
    for (int64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
        uint32_t keep = 0;                                               // bits of the word that belong to neighbouring envs
        if (w == w0 && (t0 & 31)) keep |= (1u << (t0 & 31)) - 1u;
        if (w == w1 - 1 && (t1 & 31)) keep |= ~((1u << (t1 & 31)) - 1u);
        if (keep) atomicAnd(seen + w, keep); else seen[w] = 0u;
    }
    if (threadIdx.x == 0) potential[n] = 0;
}

// one thread per (env, agent): who is in my crosshairs — the centre rays of the two middle pooled pixels (deathmatch.py:56,
// 76-79) — then damage / health (deathmatch.py:62-70). matchings[n][shooter][target]; hits = targets hit, wounds = shooters.
__global__ void __launch_bounds__(128) shoot_kernel(const int32_t* __restrict__ indices, const float* __restrict__ positions,
                                                    const float* __restrict__ bounds, uint8_t* matchings, float* hits, float* health,
                                                    float* damage, int32_t n_envs, int32_t A, int32_t R, int32_t sub, int32_t F,
                                                    float clearance) {
    const int n = blockIdx.x;
    extern __shared__ int wounds[];                                    // [A]
    for (int a = threadIdx.x; a < A; a += blockDim.x) wounds[a] = 0;
    __syncthreads();
    const int Ro = R / sub;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const int32_t* row = indices + ((int64_t)n * A + a) * R;
        int h = 0;
        uint8_t* mrow = matchings + ((int64_t)n * A + a) * A;
        for (int t = 0; t < A; t++) mrow[t] = 0;
        for (int px = Ro / 2 - 1; px <= Ro / 2; px++) {
            if (px < 0 || px >= Ro) continue;
            const int line = __ldg(row + px * sub + sub / 2);
            const int who = line >= 0 ? line / F : -1;
            if (who >= 0 && who < A && !mrow[who]) { mrow[who] = 1; h++; atomicAdd(&wounds[who], 1); }
        }
        hits[(int64_t)n * A + a] = (float)h;
        damage[(int64_t)n * A + a] = __fadd_rn(damage[(int64_t)n * A + a], __fmul_rn(.05f, (float)h));      // (torch: two roundings)
    }
    __syncthreads();
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        const int64_t i = (int64_t)n * A + a;
        const float px = positions[2 * i], py = positions[2 * i + 1];
        // bounds are (height, width) of the mask grid in metres, compared against (x, y) as the reference does (deathmatch.py:66)
        const bool outside = px < -clearance || py < -clearance || px > bounds[2 * n] + clearance || py > bounds[2 * n + 1] + clearance;
        health[i] = __fadd_rn(health[i], __fsub_rn(__fmul_rn(-.05f, (float)wounds[a] + (outside ? 1.f : 0.f)), .001f));
    }
}

__device__ __forceinline__ uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

// one thread per (env, agent): flagged agents go to one of their precomputed spawn points, velocities zeroed
__global__ void __launch_bounds__(256) respawn_kernel(const uint8_t* __restrict__ reset, const float* __restrict__ spawn_positions,
                                                      const float* __restrict__ spawn_angles, msb_agents ag, int64_t n_agents_total,
                                                      int32_t n_spawns, uint32_t seed, uint32_t tick, int32_t* choices) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_agents_total || !reset[i]) return;
    const int c = choices ? choices[i] % n_spawns : (int)(hash3(seed, tick, (uint32_t)i) % (uint32_t)n_spawns);
    ag.angles[i] = spawn_angles[i * n_spawns + c];
    ag.positions[2 * i] = spawn_positions[2 * (i * n_spawns + c)];
    ag.positions[2 * i + 1] = spawn_positions[2 * (i * n_spawns + c) + 1];
    ag.velocity[2 * i] = 0.f; ag.velocity[2 * i + 1] = 0.f;
    ag.angvelocity[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// pack_obs_kernel: the three observation heads of a batch into one row per env, [rgb | d | imu], in one pass — the send
// buffer of ShardedCore's all-gather (sharding.PackedObs) — as fp32, fp16, or 8-bit images with an fp16 imu.
// Same values as the PyTorch form: x -> half by round-to-nearest-even; x -> uint8 as
// clamp(round_half_even(255 x), 0, 255).
// ---------------------------------------------------------------------------------------------------------------
#include <cuda_fp16.h>
template <typename T> struct Pack4 { T v[4]; };

__device__ __forceinline__ unsigned char quant8(float x) { return (unsigned char)fminf(fmaxf(rintf(__fmul_rn(x, 255.f)), 0.f), 255.f); }

// grid (ceil(units / 256), N): blockIdx.y = env, no division anywhere. VEC: the image part moves four pixels per thread
// (float4 in, 4 / 8 / 16 bytes out) — needs 4 | A * 3 * ro and 4 | A * ro; the imu (3 A values) goes one per thread.
template <bool VEC>
__global__ void __launch_bounds__(256) pack_obs_kernel(const float* __restrict__ rgb, const float* __restrict__ d,
                                                       const float* __restrict__ imu, unsigned char* rows, int n_img_rgb, int n_img_d,
                                                       int n_imu, int64_t row_bytes, int mode, int imu_off) {
    const int64_t n = blockIdx.y;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned char* row = rows + n * row_bytes;
    const int n_img = n_img_rgb + n_img_d;
    const int W = VEC ? 4 : 1;
    const int img_units = n_img / W;
    if (u < img_units) {
        const int e = u * W;
        const float* src = e < n_img_rgb ? rgb + n * n_img_rgb + e : d + n * n_img_d + (e - n_img_rgb);
        float x[4];
        if (VEC) { const float4 v = __ldg(reinterpret_cast<const float4*>(src)); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
        else x[0] = __ldg(src);
        if (mode == 0) {
            if (VEC) *reinterpret_cast<float4*>(row + 4 * (size_t)e) = make_float4(x[0], x[1], x[2], x[3]);
            else reinterpret_cast<float*>(row)[e] = x[0];
        } else if (mode == 1) {
            if (VEC) {
                Pack4<__half> h;
#pragma unroll
                for (int i = 0; i < 4; i++) h.v[i] = __float2half_rn(x[i]);
                *reinterpret_cast<uint2*>(row + 2 * (size_t)e) = *reinterpret_cast<uint2*>(&h);
            } else reinterpret_cast<__half*>(row)[e] = __float2half_rn(x[0]);
        } else {
            if (VEC) {
                Pack4<unsigned char> q;
#pragma unroll
                for (int i = 0; i < 4; i++) q.v[i] = quant8(x[i]);
                *reinterpret_cast<uint32_t*>(row + e) = *reinterpret_cast<uint32_t*>(&q);
            } else row[e] = quant8(x[0]);
        }
    } else if (u < img_units + n_imu) {
        const int j = u - img_units;
        const float x = __ldg(imu + n * n_imu + j);
        if (mode == 0) reinterpret_cast<float*>(row)[n_img + j] = x;
        else if (mode == 1) reinterpret_cast<__half*>(row)[n_img + j] = __float2half_rn(x);
        else reinterpret_cast<__half*>(row + imu_off)[j] = __float2half_rn(x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: C ABI
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static long long g_opt_nch = 0;          // 0 = auto
static long long g_opt_threads = 0;      // 0 = auto
static long long g_opt_skip_dyn = 0;     // debug
static long long g_opt_no_vis = 0;       // tests: ignore the visibility grid
static long long g_opt_idx64 = 0;        // tests: 64-bit output indexing even when 32 bits would do
static long long g_opt_stage_rec = 0;    // 0: auto (when n_agents * res >= 256), 1: always, 2: never
static long long g_opt_no_env_order = 0; // A/B: CTA b takes env b
static long long g_opt_no_prefetch = 0;  // A/B: physics_kernel does not prefetch view_kernel's tables
static long long g_opt_dyn_window = 0;   // 0 = default (DYN_MIN_WINDOW); 1, 2, 4, 8: pixels per queue entry when subsample is smaller
static long long g_opt_dyn_warps = 0;    // warps sharing one queue entry in dyn_kernel (default 2)
static long long g_opt_fused_step = 0;   // 1: msb_step runs physics inside view_kernel (measured slower: the physics latency
                                         // chain adds to every CTA's life instead of running at its own high occupancy)
static long long g_opt_bake_brute = 0;  // 1: bake tests every static line per (texel, light) like the reference, even when the spatial table exists
static long long g_opt_view_ctas_per_sm = 0;   // view_kernel: cap the resident CTAs per SM (by padding its shared memory); 0: auto
static long long g_opt_persist = 0;      // 1: render with tick_kernel (persistent grid) instead of view_kernel + dyn_kernel
static long long g_opt_merge_dyn = 0;    // tick_kernel lights the agent-hit windows itself — 0 / 1: yes, 2: no (dyn_kernel follows)
static long long g_opt_dyn_groups = 0;   // merged second pass: tickets per queue entry (1, 2 or 4; default 4)
static long long g_opt_stages = 0;       // tick_kernel: envs staged per CTA (2..4; 0: auto)
static long long g_opt_no_sched = 0;     // tick_kernel: deal the envs round-robin instead of first come, first served
static unsigned long long* g_stats = nullptr;   // device counters, enabled by option "stats"

// Optional per-kernel timing with CUDA events on the launching stream (option "timing" = 1): bench.py uses it to
// report each kernel's live share of the step. Kinds: 0 physics, 1 render, 2 (unused), 3 dyn, 4 fused step, 5 bake.
enum { TK_PHYSICS = 0, TK_RENDER = 1, TK_SHADE = 2, TK_DYN = 3, TK_STEP = 4, TK_BAKE = 5, TK_KINDS = 6 };
static long long g_opt_timing = 0;
static const int TIMING_RING = 4096;
static cudaEvent_t g_ev[2 * TIMING_RING];
static int g_ev_kind[TIMING_RING];
static int g_ev_n = 0;
static bool g_ev_ready = false;
static double g_time_ms[TK_KINDS] = {0, 0, 0, 0, 0, 0};
static long long g_time_n[TK_KINDS] = {0, 0, 0, 0, 0, 0};

static void timing_flush() {
    for (int i = 0; i < g_ev_n; i++) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_ev[2 * i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, g_ev[2 * i], g_ev[2 * i + 1]) == cudaSuccess) {
            g_time_ms[g_ev_kind[i]] += ms;
            g_time_n[g_ev_kind[i]]++;
        }
    }
    g_ev_n = 0;
}
struct TimedLaunch {
    int slot;
    cudaStream_t st;
    TimedLaunch(int kind, cudaStream_t st_) : slot(-1), st(st_) {
        if (!g_opt_timing) return;
        if (!g_ev_ready) {
            for (int i = 0; i < 2 * TIMING_RING; i++) cudaEventCreate(&g_ev[i]);
            g_ev_ready = true;
        }
        if (g_ev_n == TIMING_RING) timing_flush();
        slot = g_ev_n++;
        g_ev_kind[slot] = kind;
        cudaEventRecord(g_ev[2 * slot], st);
    }
    ~TimedLaunch() { if (slot >= 0) cudaEventRecord(g_ev[2 * slot + 1], st); }
};

// Kernel launch, optionally as a programmatic dependent of the kernel ahead in the stream (see msb_math.cuh).
static long long g_opt_pdl = 1;
static cudaError_t launch(void (*fn)(KArgs), int grid, int block, size_t smem, cudaStream_t st, bool dependent, const KArgs& k) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && g_opt_pdl && !g_opt_timing) ? 1 : 0;   // per-kernel timing wants the kernels apart
    return cudaLaunchKernelEx(&cfg, fn, k);
}

static int fail(const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return 1;
}
static int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return 1;
}

extern "C" int msb_abi_version(void) { return MSB_ABI_VERSION; }
extern "C" const char* msb_last_error(void) { return g_err; }
extern "C" int64_t msb_launch_count(void) { return g_launches.load(); }

extern "C" int msb_params_init(msb_params* p, float agent_radius, int32_t res, float fov, float fps) {
    if (!p) return fail("%s", "msb_params_init: null params");
    if (!(fov < 180.f) || !(fov > 0.f)) return fail("%s", "msb_params_init: fov must be in (0, 180)");
    if (res < 1) return fail("%s", "msb_params_init: res must be >= 1");
    if (!(fps > 0.f)) return fail("%s", "msb_params_init: fps must be positive");
    memset(p, 0, sizeof(*p));
    p->res = res;
    p->agent_radius = agent_radius;
    p->half_screen = tanf(CUDART_PI_F / 180.f * fov / 2.);   // exactly kernels.cu:22
    p->fps = fps;
    p->fov = fov;
    return 0;
}

extern "C" int msb_set_option(const char* name, int64_t value) {
    if (!strcmp(name, "nch")) { g_opt_nch = value; return 0; }
    if (!strcmp(name, "threads")) { g_opt_threads = value; return 0; }
    if (!strcmp(name, "debug_skip_dyn")) { g_opt_skip_dyn = value; return 0; }
    if (!strcmp(name, "fused_step")) { g_opt_fused_step = value; return 0; }
    if (!strcmp(name, "dyn_window")) { g_opt_dyn_window = value; return 0; }
    if (!strcmp(name, "dyn_warps")) { g_opt_dyn_warps = value; return 0; }
    if (!strcmp(name, "pdl")) { g_opt_pdl = value; return 0; }
    if (!strcmp(name, "no_vis")) { g_opt_no_vis = value; return 0; }
    if (!strcmp(name, "no_prefetch")) { g_opt_no_prefetch = value; return 0; }
    if (!strcmp(name, "no_env_order")) { g_opt_no_env_order = value; return 0; }
    if (!strcmp(name, "stage_rec")) { g_opt_stage_rec = value; return 0; }
    if (!strcmp(name, "idx64")) { g_opt_idx64 = value; return 0; }
    if (!strcmp(name, "persist")) { g_opt_persist = value; return 0; }
    if (!strcmp(name, "view_ctas_per_sm")) { g_opt_view_ctas_per_sm = value; return 0; }
    if (!strcmp(name, "bake_brute")) { g_opt_bake_brute = value; return 0; }
    if (!strcmp(name, "merge_dyn")) { g_opt_merge_dyn = value; return 0; }
    if (!strcmp(name, "dyn_groups")) { g_opt_dyn_groups = value; return 0; }
    if (!strcmp(name, "stages")) { g_opt_stages = value; return 0; }
    if (!strcmp(name, "no_sched")) { g_opt_no_sched = value; return 0; }
    if (!strcmp(name, "timing")) {
        timing_flush();
        g_opt_timing = value;
        for (int i = 0; i < TK_KINDS; i++) { g_time_ms[i] = 0; g_time_n[i] = 0; }
        return 0;
    }
    if (!strcmp(name, "stats")) {
        if (value && !g_stats) {
            if (check(cudaMalloc(&g_stats, STAT_SLOTS * sizeof(unsigned long long)), "cudaMalloc(stats)")) return 1;
            return check(cudaMemset(g_stats, 0, STAT_SLOTS * sizeof(unsigned long long)), "cudaMemset(stats)");
        }
        if (!value && g_stats) { cudaFree(g_stats); g_stats = nullptr; }
        return 0;
    }
    if (!strcmp(name, "stats_reset")) {
        if (g_stats) return check(cudaMemset(g_stats, 0, STAT_SLOTS * sizeof(unsigned long long)), "cudaMemset(stats)");
        return 0;
    }
    return fail("msb_set_option: unknown option '%s'", name);
}

extern "C" int64_t msb_get_option(const char* name) {
    // "time_ns_<kind>" / "time_count_<kind>": accumulated device time of that kernel since "timing" was set
    static const char* kinds[TK_KINDS] = {"physics", "render", "shade", "dyn", "step", "bake"};
    if (!strncmp(name, "time_", 5)) {
        timing_flush();
        for (int i = 0; i < TK_KINDS; i++) {
            char a[64], b[64];
            snprintf(a, sizeof(a), "time_ns_%s", kinds[i]);
            snprintf(b, sizeof(b), "time_count_%s", kinds[i]);
            if (!strcmp(name, a)) return (int64_t)(g_time_ms[i] * 1e6);
            if (!strcmp(name, b)) return (int64_t)g_time_n[i];
        }
        return -1;
    }
    if (!strcmp(name, "nch")) return g_opt_nch;
    if (!strcmp(name, "threads")) return g_opt_threads;
    if (!strncmp(name, "stat", 4) && g_stats) {
        unsigned long long h[STAT_SLOTS];
        if (cudaMemcpy(h, g_stats, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (name[5] >= '0' && name[5] <= '9') { const int i = atoi(name + 5); return i < STAT_SLOTS ? (int64_t)h[i] : -1; }
        if (!strcmp(name, "stat_tests")) return (int64_t)h[STAT_TESTS];
        if (!strcmp(name, "stat_groups")) return (int64_t)h[STAT_GROUPS];
        if (!strcmp(name, "stat_dyn_rays")) return (int64_t)h[STAT_DYN_RAYS];
        if (!strcmp(name, "stat_dyn_iters")) return (int64_t)h[STAT_DYN_ITERS];
        if (!strcmp(name, "stat_dyn_entries")) return (int64_t)h[STAT_DYN_ENTRIES];
        if (!strcmp(name, "stat_replays")) return (int64_t)h[STAT_REPLAYS];
        if (!strcmp(name, "stat_dyn_cycles")) return (int64_t)h[STAT_DYN_CYCLES];
        if (!strcmp(name, "stat_dyn_maxcyc")) return (int64_t)h[STAT_DYN_MAXCYC];
        if (!strcmp(name, "stat_dyn_warpmax")) return (int64_t)h[STAT_DYN_WARPMAX];
        if (!strcmp(name, "stat_dyn_slow")) return (int64_t)h[STAT_DYN_SLOW];
        if (!strcmp(name, "stat_dyn_kernel")) return (int64_t)h[STAT_DYN_KERNEL];
        if (!strcmp(name, "stat_dyn_scans")) return (int64_t)h[STAT_DYN_SCANS];
        if (!strcmp(name, "stat_dyn_scans_lit")) return (int64_t)h[STAT_DYN_SCANS_LIT];
        if (!strcmp(name, "stat_dyn_iters_lit")) return (int64_t)h[STAT_DYN_ITERS_LIT];
    }
    return -1;
}

static int validate(const msb_params* p, const msb_scenery* s, bool need_table) {
    if (!p || !s) return fail("%s", "null params/scenery");
    if (s->n_envs < 0 || s->n_agents < 1 || s->n_model < 0) return fail("%s", "bad scenery dimensions");
    if (s->max_lines < s->n_agents * s->n_model) return fail("%s", "scenery.max_lines is smaller than n_agents*n_model");
    if (s->max_lines > 14000) return fail("%s", "scene too large: more than 14000 segments in one environment");
    if (need_table && s->n_envs > 0 &&
        (!s->occ_lines || !s->occ_boxes || !s->box_starts || !s->occ_rec || !s->occ_meta || s->occ_run != VRUN))
        return fail("%s", "scenery has no spatial table (occ_lines, occ_boxes, box_starts, occ_rec, occ_meta with occ_run = 16): "
                          "build it once per scenery, see include/megastep_b200.h");
    return 0;
}

static void fill(KArgs& k, const msb_params* p, const msb_scenery* s, const msb_agents* a) {
    memset(&k, 0, sizeof(k));
    k.p = *p;
    k.s = *s;
    if (a) k.a = *a;
    k.seg_cap = s->max_lines > 0 ? s->max_lines : 1;
    const int w = s->max_lines - s->n_agents * s->n_model;
    k.wcap = w > 0 ? ((w + VRUN - 1) / VRUN) * VRUN : VRUN;
    k.inv_fps = 1.0f / p->fps;
    k.ray_blocks = 1;
    k.stats = g_stats;
    k.debug_skip_dyn = (int32_t)g_opt_skip_dyn;
    k.debug_no_vis = (int32_t)g_opt_no_vis;
    k.env_order = g_opt_no_env_order ? nullptr : s->env_order;
    k.xclip = 0.5f * p->agent_radius / sqrtf(1.f + p->half_screen * p->half_screen);
}

static void set_movement(KArgs& k, const msb_params* p, const msb_movement* mv) {
    if (!mv) return;
    k.mv = *mv;
    k.has_mv = 1;
    k.mv_keep = (float)(1.0 - (double)mv->decay);
    k.mv_dv = (float)((double)mv->accel / (double)p->fps);
    k.mv_dw = (float)((double)mv->ang_accel / (double)p->fps);
}

static int launch_physics(const KArgs& k, cudaStream_t st) {
    int warps = k.s.n_agents < 4 ? k.s.n_agents : 4;
    if (g_opt_threads >= 32 && g_opt_threads <= 128) warps = (int)(g_opt_threads / 32);
    const size_t sm = (size_t)k.s.n_agents * ST_STRIDE * 4 * 2;
    if (sm > 48 * 1024) return fail("%s", "too many agents per environment");
    {
        TimedLaunch timed(TK_PHYSICS, st);
        launch(physics_kernel, k.s.n_envs, 32 * warps, sm, st, false, k);
    }
    g_launches++;
    return check(cudaGetLastError(), "physics_kernel launch");
}

// how the rays of one agent are split over warps: NCH 32-ray chunks per warp, RB warps per agent
static void plan_view(const msb_params* p, const msb_scenery* s, int* nch, int* rb, int* threads) {
    const int chunks = (p->res + 31) / 32;
    int n = 2;      // two chunks per warp: measured best (four: more registers live, spills; one: every warp re-bins)
    if (g_opt_nch == 1 || g_opt_nch == 2 || g_opt_nch == 4) n = (int)g_opt_nch;
    while (n > 1 && n > chunks) n >>= 1;
    if (!g_opt_nch) while (n > 1 && s->n_agents * ((chunks + n - 1) / n) < 4) n >>= 1;   // keep a CTA at 4 warps of work
    *nch = n;
    *rb = (chunks + n - 1) / n;
    int t = 32 * s->n_agents * (*rb);
    if (t > 128) t = 128;       // more items than warps: a warp takes several, one after the other
    if (g_opt_threads >= 32 && g_opt_threads <= 256) t = (int)(g_opt_threads / 32) * 32;
    *threads = t;
}

static int launch_view(KArgs& k, bool phys, int nch, int threads, cudaStream_t st) {
    k.stage_rec = g_opt_stage_rec ? (g_opt_stage_rec == 1) : (k.s.n_agents * k.p.res >= 256);
    k.idx32 = (!g_opt_idx64 && 3ll * k.s.n_envs * k.s.n_agents * (long long)k.p.res < (1ll << 32)) ? 1 : 0;
    k.out_mask = (k.out.indices ? OUT_INDICES : 0) | (k.out.locations ? OUT_LOCATIONS : 0) | (k.out.dots ? OUT_DOTS : 0) |
                 (k.out.distances ? OUT_DISTANCES : 0) | (k.out.screen ? OUT_SCREEN : 0) |
                 (k.has_obs && k.obs.rgb ? OUT_RGB : 0) | (k.has_obs && k.obs.depth ? OUT_DEPTH : 0) | (k.has_obs && k.obs.imu ? OUT_IMU : 0);
    size_t sm = vsmem_bytes(k.wcap, threads / 32, k.s.n_agents, k.s.n_agents * k.s.n_model, k.stage_rec != 0);
    if (sm > 227 * 1024 && k.stage_rec) {                      // a very large env: leave the records in global memory
        k.stage_rec = 0;
        sm = vsmem_bytes(k.wcap, threads / 32, k.s.n_agents, k.s.n_agents * k.s.n_model, false);
    }
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
    if (g_opt_view_ctas_per_sm >= 1 && g_opt_view_ctas_per_sm <= 16) {
        // fewer resident CTAs than the registers allow: the grid's last, partial wave is what it costs to have more
        const size_t per = (size_t)(227 * 1024) / (size_t)g_opt_view_ctas_per_sm - 1024;
        if (per > sm) sm = per & ~size_t(15);
    }
#define MSB_LAUNCH(N)                                                                                            \
    {                                                                                                            \
        auto fn = phys ? (k.stats ? view_kernel<N, true, true> : view_kernel<N, true, false>)                    \
                       : (k.stats ? view_kernel<N, false, true> : view_kernel<N, false, false>);                 \
        if (sm > 48 * 1024 && check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), \
                                    "cudaFuncSetAttribute"))                                                     \
            return 1;                                                                                            \
        launch(fn, k.s.n_envs, threads, sm, st, true, k);                                                        \
    }
    {
        TimedLaunch timed(phys ? TK_STEP : TK_RENDER, st);
        switch (nch) {
            case 1: MSB_LAUNCH(1); break;
            case 2: MSB_LAUNCH(2); break;
            default: MSB_LAUNCH(4); break;
        }
    }
#undef MSB_LAUNCH
    g_launches++;
    return check(cudaGetLastError(), "view_kernel launch");
}

static int dyn_window(int sub) {
    int w = DYN_MIN_WINDOW;
    if (g_opt_dyn_window == 1 || g_opt_dyn_window == 2 || g_opt_dyn_window == 4 || g_opt_dyn_window == 8) w = (int)g_opt_dyn_window;
    return sub > w ? sub : w;
}

// workspace layout: int ctrl[4] (queue) | int sched[4] (tick_kernel's env counter) | occluder cache int[N][A][32] |
// queue entries of DYN_HDR + 32 * PS bytes
enum { WS_HEAD = 32 };
static int64_t cache_bytes(const msb_scenery* s) { return (int64_t)s->n_envs * s->n_agents * 32 * 4; }

static int set_workspace(KArgs& k, const msb_workspace* ws) {
    k.dyn_ctrl = nullptr;
    k.dyn_entries = nullptr;
    k.dyn_cache = nullptr;
    k.dyn_cap = 0;
    k.sched = nullptr;
    if (!ws || !ws->ptr) return 0;
    if (((uintptr_t)ws->ptr & 15) != 0) return fail("%s", "workspace must be 16-byte aligned");
    if (ws->bytes < WS_HEAD) return 0;
    k.dyn_ctrl = reinterpret_cast<int*>(ws->ptr);
    k.sched = g_opt_no_sched ? nullptr : reinterpret_cast<int*>(ws->ptr) + 4;
    if (k.s.n_agents == 1) return 0;    // a lone agent can only ever hit its own model (if at all): no second pass to launch
    const int sub = k.has_obs ? k.obs.subsample : 1;
    k.dyn_window = dyn_window(sub);
    const int64_t head = WS_HEAD + cache_bytes(&k.s);
    const int64_t cap = (ws->bytes - head) / (DYN_HDR + 32 * k.dyn_window);
    if (cap < 1) return 0;
    k.dyn_cache = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws->ptr) + WS_HEAD);
    k.dyn_entries = reinterpret_cast<unsigned char*>(ws->ptr) + head;
    k.dyn_cap = cap > 0x7fffffff ? 0x7fffffff : (int32_t)cap;
    return 0;
}

// resident CTAs per SM of a kernel at a given shape, per device (a process may drive several GPUs)
static int occupancy(const void* fn, int threads, size_t smem, int* sms_out) {
    struct Key { const void* fn; int dev, threads; size_t smem; int per_sm, sms; };
    static std::mutex mu;
    static Key cache[64];
    static int n_cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_cached; i++)
        if (cache[i].fn == fn && cache[i].dev == dev && cache[i].threads == threads && cache[i].smem == smem) {
            if (sms_out) *sms_out = cache[i].sms;
            return cache[i].per_sm;
        }
    Key key = {fn, dev, threads, smem, 0, 0};
    cudaDeviceGetAttribute(&key.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&key.per_sm, fn, threads, smem);
    if (key.per_sm < 1) key.per_sm = 1;
    if (n_cached < 64) cache[n_cached++] = key;
    if (sms_out) *sms_out = key.sms;
    return key.per_sm;
}

static int launch_dyn(const KArgs& k, cudaStream_t st) {
    if (!k.dyn_entries) return 0;
    {
        TimedLaunch timed(TK_DYN, st);
        const int threads = (g_opt_dyn_warps >= 1 && g_opt_dyn_warps <= 4) ? 32 * (int)g_opt_dyn_warps : 64;   // measured: 2 warps per entry
        const int scale = 128 / threads;            // same number of resident threads whatever the CTA size
        int sms = 0;
        if (k.stats) {
            const int per_sm = occupancy((const void*)dyn_kernel<true>, 128, 0, &sms);       // exactly one wave
            launch(dyn_kernel<true>, sms * per_sm * scale, threads, 0, st, true, k);
        } else {
            const int per_sm = occupancy((const void*)dyn_kernel<false>, 128, 0, &sms);
            launch(dyn_kernel<false>, sms * per_sm * scale, threads, 0, st, true, k);
        }
    }
    g_launches++;
    return check(cudaGetLastError(), "dyn_kernel launch");
}

// tick_kernel: the persistent form of view_kernel (+ dyn_kernel). *merged: the second pass ran inside it.
static bool want_tick(const KArgs& k, int rb) {
    // Opt-in (option "persist" = 1). Measured on the benchmark (4096 envs x 4 agents x 128 rays, B200): 112-121 us against
    // view_kernel's 102 — with ~7 envs per persistent CTA the pipeline's look-ahead (staged envs) costs more in the grid's
    // tail than the barriers it removes; see DESIGN.md.
    (void)rb;
    if (g_opt_fused_step) return false;
    return g_opt_persist == 1;
}

static int launch_tick(KArgs& k, int nch, cudaStream_t st, bool* merged) {
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    k.idx32 = (!g_opt_idx64 && 3ll * k.s.n_envs * k.s.n_agents * (long long)k.p.res < (1ll << 32)) ? 1 : 0;
    k.out_mask = (k.out.indices ? OUT_INDICES : 0) | (k.out.locations ? OUT_LOCATIONS : 0) | (k.out.dots ? OUT_DOTS : 0) |
                 (k.out.distances ? OUT_DISTANCES : 0) | (k.out.screen ? OUT_SCREEN : 0) |
                 (k.has_obs && k.obs.rgb ? OUT_RGB : 0) | (k.has_obs && k.obs.depth ? OUT_DEPTH : 0) | (k.has_obs && k.obs.imu ? OUT_IMU : 0);
    k.p_items = A * k.ray_blocks;
    int threads = 256;
    if (g_opt_threads >= 64 && g_opt_threads <= 1024) threads = (int)(g_opt_threads / 32) * 32;
    const int bound = threads <= 256 ? 256 : (threads <= 512 ? 512 : 1024);     // the instantiation: bound x (1024 / bound) CTAs fill an SM's registers
    const int per_sm_target = 1024 / bound;
    const int nwarps = threads / 32;
    const size_t scr = (size_t)nwarps * (32 * nch < 64 ? 64 : 32 * nch) * 16;
    const size_t ctl = (sizeof(PCtl) + 15) & ~size_t(15);
    // stages and whether the rows' records are staged too: as much as keeps the SM's registers in use
    const size_t budget = (227 * 1024 - per_sm_target * 1024) / per_sm_target;
    bool rec = g_opt_stage_rec != 2;
    auto total = [&](int stages, bool r) { return ctl + stages * pstage_bytes(k.wcap, A, AF, r) + scr; };
    // enough stages for every warp to have an item, plus one being refilled
    int S = (nwarps + k.p_items - 1) / k.p_items + 1;
    if (S < 2) S = 2;
    if (g_opt_stages >= 2 && g_opt_stages <= PMAXS) S = (int)g_opt_stages;
    if (S > PMAXS) S = PMAXS;
    if (rec && g_opt_stage_rec != 1 && total(S, true) > budget) rec = false;
    while (S > 2 && !g_opt_stages && total(S, rec) > budget) S--;
    if (S > nwarps) S = nwarps;
    if (total(S, rec) > 227 * 1024 && rec) rec = false;
    while (S > 2 && total(S, rec) > 227 * 1024) S--;
    if (total(S, rec) > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
    k.stage_rec = rec ? 1 : 0;
    k.p_stages = S;
    k.p_stage_bytes = (int32_t)pstage_bytes(k.wcap, A, AF, rec);
    k.dyn_groups = (g_opt_dyn_groups == 1 || g_opt_dyn_groups == 2 || g_opt_dyn_groups == 4) ? (int)g_opt_dyn_groups : 4;
    const size_t sm = total(S, rec);
    const bool merge = g_opt_merge_dyn != 2 && k.dyn_entries != nullptr && k.sched != nullptr;
    *merged = merge;
#define MSB_GO(fn)                                                                                               \
    {                                                                                                            \
        if (sm > 48 * 1024 && check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), \
                                    "cudaFuncSetAttribute"))                                                     \
            return 1;                                                                                            \
        int sms = 0;                                                                                             \
        const int per_sm = occupancy((const void*)fn, threads, sm, &sms);                                       \
        int grid = sms * per_sm;                                                                                 \
        if (grid > k.s.n_envs) grid = k.s.n_envs;                                                               \
        launch(fn, grid, threads, sm, st, true, k);                                                              \
    }
#define MSB_LAUNCH(N)                                                                                            \
    {                                                                                                            \
        if (bound == 256) {                                                                                      \
            if (merge) { if (k.stats) MSB_GO((tick_kernel<N, true, true, 256>)) else MSB_GO((tick_kernel<N, true, false, 256>)) }   \
            else { if (k.stats) MSB_GO((tick_kernel<N, false, true, 256>)) else MSB_GO((tick_kernel<N, false, false, 256>)) }       \
        } else if (bound == 512) {                                                                               \
            if (merge) MSB_GO((tick_kernel<N, true, false, 512>)) else MSB_GO((tick_kernel<N, false, false, 512>))                  \
        } else {                                                                                                 \
            if (merge) { if (k.stats) MSB_GO((tick_kernel<N, true, true, 1024>)) else MSB_GO((tick_kernel<N, true, false, 1024>)) } \
            else { if (k.stats) MSB_GO((tick_kernel<N, false, true, 1024>)) else MSB_GO((tick_kernel<N, false, false, 1024>)) }     \
        }                                                                                                        \
    }
    {
        TimedLaunch timed(TK_RENDER, st);
        switch (nch) {
            case 1: MSB_LAUNCH(1); break;
            case 2: MSB_LAUNCH(2); break;
            default: MSB_LAUNCH(4); break;
        }
    }
#undef MSB_LAUNCH
#undef MSB_GO
    g_launches++;
    return check(cudaGetLastError(), "tick_kernel launch");
}

// render (+ heads) after the agents have been moved: the persistent kernel where an env has enough items, else one CTA per env
static int launch_render(KArgs& k, int nch, int rb, int threads, cudaStream_t st) {
    k.view_ctas = k.s.n_envs;       // what the second pass counts up to: envs whose rays are done
    if (want_tick(k, rb)) {
        bool merged = false;
        if (launch_tick(k, nch, st, &merged)) return 1;
        return merged ? 0 : launch_dyn(k, st);
    }
    if (launch_view(k, false, nch, threads, st)) return 1;
    return launch_dyn(k, st);
}

extern "C" int64_t msb_workspace_bytes(const msb_params* p, const msb_scenery* s, int32_t subsample) {
    if (!p || !s || subsample < 1) return 0;
    const int PS = subsample > 8 ? subsample : 8;          // sized for the largest window the options allow
    const int64_t windows = (int64_t)s->n_envs * s->n_agents * ((p->res + PS - 1) / PS);
    // room for a quarter of all pixel windows to contain an agent-hit ray (overflow falls back to inline, still exact)
    int64_t cap = windows / 4;
    if (cap < 16384) cap = windows < 16384 ? windows : 16384;
    return WS_HEAD + cache_bytes(s) + cap * (DYN_HDR + 32 * PS);
}

static void set_obs(KArgs& k, const msb_obs_out* obs) {
    if (!obs) return;
    k.obs = *obs;
    k.has_obs = 1;
    k.inv_max_depth = 1.0f / obs->max_depth;
    k.inv_speed = 1.0f / obs->speed_scale;
    k.inv_ang = 1.0f / obs->ang_scale;
    k.inv_sub = 1.0f / (float)obs->subsample;
    k.sub_shift = __builtin_ctz((unsigned)obs->subsample);
}

static int check_obs(const msb_params* p, const msb_obs_out* obs) {
    if (!obs) return 0;
    const int sub = obs->subsample;
    if (sub < 1 || sub > 32 || (sub & (sub - 1)) || p->res % sub) return fail("%s", "obs.subsample must be a power of two <= 32 dividing res");
    return 0;
}

extern "C" int msb_physics(const msb_params* p, const msb_scenery* s, const msb_agents* a, float* progress,
                           void* cuda_stream) {
    if (validate(p, s, true)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_physics: null agents");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    k.progress = progress;
    return launch_physics(k, (cudaStream_t)cuda_stream);
}

extern "C" int msb_move(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv, float* progress,
                        void* cuda_stream) {
    if (validate(p, s, true)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_move: null agents");
    if (!mv || !mv->actions) return fail("%s", "msb_move: movement without actions");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    k.progress = progress;
    set_movement(k, p, mv);
    return launch_physics(k, (cudaStream_t)cuda_stream);
}

extern "C" int msb_render(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_render_out* out,
                          const msb_obs_out* obs, const msb_workspace* ws, void* cuda_stream) {
    if (validate(p, s, true) || check_obs(p, obs)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_render: null agents");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    if (out) k.out = *out;
    set_obs(k, obs);
    if (set_workspace(k, ws)) return 1;
    int nch, rb, threads;
    plan_view(p, s, &nch, &rb, &threads);
    k.ray_blocks = rb;
    k.rb_shift = (rb & (rb - 1)) ? -1 : __builtin_ctz((unsigned)rb);
    return launch_render(k, nch, rb, threads, (cudaStream_t)cuda_stream);
}

extern "C" int msb_step(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv,
                        float* progress, const msb_render_out* out, const msb_obs_out* obs, const msb_workspace* ws,
                        void* cuda_stream) {
    if (validate(p, s, true) || check_obs(p, obs)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_step: null agents");
    if (mv && !mv->actions) return fail("%s", "msb_step: movement without actions");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    k.progress = progress;
    if (out) k.out = *out;
    set_obs(k, obs);
    set_movement(k, p, mv);
    if (set_workspace(k, ws)) return 1;
    int nch, rb, threads;
    plan_view(p, s, &nch, &rb, &threads);
    k.ray_blocks = rb;
    k.rb_shift = (rb & (rb - 1)) ? -1 : __builtin_ctz((unsigned)rb);
    if (g_opt_fused_step) {
        k.view_ctas = k.s.n_envs;
        if (launch_view(k, true, nch, threads, (cudaStream_t)cuda_stream)) return 1;
        return launch_dyn(k, (cudaStream_t)cuda_stream);
    }
    // physics (with the movement prologue) at its own, higher occupancy; then render with the heads. One CTA per env
    // stages its table right at its start: physics sends it on its way from HBM to L2; the persistent kernel stages an
    // env ahead, which hides that latency by itself.
    k.prefetch_view = (g_opt_no_prefetch || want_tick(k, rb)) ? 0 : 1;
    if (launch_physics(k, (cudaStream_t)cuda_stream)) return 1;
    k.prefetch_view = 0;
    return launch_render(k, nch, rb, threads, (cudaStream_t)cuda_stream);
}

// ---------------------------------------------------------------------------------------------------------------
// A whole host-to-host tick as one call: the step's three launches captured once in a CUDA graph (programmatic
// dependencies included), replayed between the upload of the actions and the download of `progress`.
// ---------------------------------------------------------------------------------------------------------------
struct msb_graph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int32_t* actions_dev;
    float* progress_dev;
    size_t n;               // N * A
    int launches;           // kernels in the captured step
};

extern "C" int msb_step_graph_create(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv,
                                     float* progress, const msb_render_out* out, const msb_obs_out* obs,
                                     const msb_workspace* ws, msb_graph** handle) {
    if (!handle) return fail("%s", "msb_step_graph_create: null handle");
    *handle = nullptr;
    if (g_opt_timing) return fail("%s", "msb_step_graph_create: per-kernel timing is on (events cannot be captured)");
    if (!s || s->n_envs <= 0) return fail("%s", "msb_step_graph_create: empty scenery");
    cudaStream_t cs;
    if (check(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "cudaStreamCreate")) return 1;
    msb_graph* g = new msb_graph();
    g->graph = nullptr; g->exec = nullptr;
    g->actions_dev = mv ? const_cast<int32_t*>(mv->actions) : nullptr;
    g->progress_dev = progress;
    g->n = (size_t)s->n_envs * s->n_agents;
    int rc = check(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
    if (!rc) {
        const long long launches = g_launches.load();
        rc = msb_step(p, s, a, mv, progress, out, obs, ws, cs);
        g->launches = (int)(g_launches.load() - launches);      // captured, not launched
        g_launches = launches;
        cudaError_t e = cudaStreamEndCapture(cs, &g->graph);
        if (!rc) rc = check(e, "cudaStreamEndCapture");
    }
    if (!rc) rc = check(cudaGraphInstantiate(&g->exec, g->graph, 0), "cudaGraphInstantiate");
    cudaStreamDestroy(cs);
    if (rc) {
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
        return 1;
    }
    *handle = g;
    return 0;
}

extern "C" int msb_step_graph_run(msb_graph* g, const int32_t* actions_host, float* progress_host, void* cuda_stream, int32_t sync) {
    if (!g || !g->exec) return fail("%s", "msb_step_graph_run: null graph");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (actions_host) {
        if (!g->actions_dev) return fail("%s", "msb_step_graph_run: the step was captured without actions");
        if (check(cudaMemcpyAsync(g->actions_dev, actions_host, g->n * sizeof(int32_t), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(actions)")) return 1;
    }
    if (check(cudaGraphLaunch(g->exec, st), "cudaGraphLaunch")) return 1;
    g_launches += g->launches;
    if (progress_host) {
        if (!g->progress_dev) return fail("%s", "msb_step_graph_run: the step was captured without progress");
        if (check(cudaMemcpyAsync(progress_host, g->progress_dev, g->n * sizeof(float), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(progress)")) return 1;
    }
    if (sync) return check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return 0;
}

extern "C" int msb_step_graph_destroy(msb_graph* g) {
    if (!g) return 0;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return 0;
}

extern "C" int msb_build_table(const msb_scenery* s, void* cuda_stream) {
    if (!s) return fail("%s", "msb_build_table: null scenery");
    if (s->n_envs < 0 || s->n_agents < 1 || s->n_model < 0) return fail("%s", "bad scenery dimensions");
    if (s->max_lines > 14000) return fail("%s", "scene too large: more than 14000 segments in one environment");
    if (!s->occ_lines || !s->occ_boxes || !s->box_starts || !s->occ_rec || !s->occ_meta || !s->occ_starts || s->occ_run != VRUN)
        return fail("%s", "msb_build_table: the table's arrays (occ_lines, occ_rec, occ_boxes, occ_meta, occ_starts) must be "
                          "allocated and box_starts filled by the caller, with occ_run = 16");
    if (!s->lines || !s->line_widths || !s->line_starts || !s->tex_widths || !s->tex_starts)
        return fail("%s", "msb_build_table: lines / textures metadata missing");
    if (s->n_envs == 0) return 0;
    msb_params p;
    memset(&p, 0, sizeof(p));
    KArgs k;
    fill(k, &p, s, nullptr);
    int n2 = 1;
    while (n2 < s->max_lines) n2 <<= 1;
    const size_t sm = (size_t)n2 * 8;
    if (sm > 48 * 1024 &&
        check(cudaFuncSetAttribute(table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "cudaFuncSetAttribute"))
        return 1;
    table_kernel<<<s->n_envs, 256, sm, (cudaStream_t)cuda_stream>>>(k);
    g_launches++;
    return check(cudaGetLastError(), "table_kernel launch");
}

extern "C" int msb_build_visibility(const msb_scenery* s, void* cuda_stream) {
    msb_params p;
    memset(&p, 0, sizeof(p));
    p.res = 1; p.fps = 1.f;
    if (validate(&p, s, true)) return 1;
    if (!s->vis || !s->vis_starts || !s->vis_meta) return fail("%s", "msb_build_visibility: scenery has no visibility grid to fill");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, &p, s, nullptr);
    const size_t sm = (size_t)k.wcap * 16 + 16;
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
    if (sm > 48 * 1024 &&
        check(cudaFuncSetAttribute(vis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "cudaFuncSetAttribute"))
        return 1;
    {
        TimedLaunch timed(TK_BAKE, (cudaStream_t)cuda_stream);
        vis_kernel<<<s->n_envs, 256, sm, (cudaStream_t)cuda_stream>>>(k);
    }
    g_launches++;
    return check(cudaGetLastError(), "vis_kernel launch");
}

extern "C" int msb_bake(const msb_params* p, const msb_scenery* s, void* cuda_stream) {
    if (validate(p, s, false)) return 1;
    if (!s->baked) return fail("%s", "msb_bake: null baked");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, nullptr);
    const bool table = !g_opt_bake_brute && s->occ_lines && s->occ_boxes && s->box_starts && s->occ_meta && s->occ_run == VRUN;
    const size_t sm = table ? (size_t)k.wcap * 16 + (size_t)(k.wcap / VRUN) * 16 + 16 : (size_t)k.seg_cap * 16 + 16;
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
    if (sm > 48 * 1024 &&
        check(cudaFuncSetAttribute(table ? bake_table_kernel : bake_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "cudaFuncSetAttribute"))
        return 1;
    {
        TimedLaunch timed(TK_BAKE, (cudaStream_t)cuda_stream);
        if (table) bake_table_kernel<<<s->n_envs, 256, sm, (cudaStream_t)cuda_stream>>>(k);
        else bake_kernel<<<s->n_envs, 256, sm, (cudaStream_t)cuda_stream>>>(k);
    }
    g_launches++;
    return check(cudaGetLastError(), "bake launch");
}

// ---------------------------------------------------------------------------------------------------------------
// environment rules (see the kernels' banner)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int msb_env_ledger_mark(const msb_scenery* s, const int32_t* indices, const float* locations, int32_t n_agents, int32_t res,
                                   uint32_t* seen, int32_t* potential, int32_t* gained, void* cuda_stream) {
    if (!s || !indices || !locations || !seen || !potential || !gained) return fail("%s", "msb_env_ledger_mark: null argument");
    if (s->n_envs == 0) return 0;
    const int64_t rays = (int64_t)s->n_envs * n_agents * res;
    const int64_t blocks = (rays + 255) / 256;
    ledger_mark_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(indices, locations, s->line_starts, s->tex_widths, s->tex_starts,
                                                                             seen, potential, gained, rays, n_agents * res);
    g_launches++;
    return check(cudaGetLastError(), "ledger_mark_kernel launch");
}

extern "C" int msb_env_ledger_clear(const msb_scenery* s, const uint8_t* reset, uint32_t* seen, int32_t* potential, void* cuda_stream) {
    if (!s || !reset || !seen || !potential) return fail("%s", "msb_env_ledger_clear: null argument");
    if (s->n_envs == 0) return 0;
    ledger_clear_kernel<<<s->n_envs, 128, 0, (cudaStream_t)cuda_stream>>>(reset, s->line_starts, s->line_widths, s->tex_widths, s->tex_starts,
                                                                       seen, potential);
    g_launches++;
    return check(cudaGetLastError(), "ledger_clear_kernel launch");
}

extern "C" int msb_env_shoot(const msb_scenery* s, const msb_agents* a, const int32_t* indices, int32_t res, int32_t subsample,
                             const float* bounds, float clearance, uint8_t* matchings, float* hits, float* health, float* damage,
                             void* cuda_stream) {
    if (!s || !a || !indices || !bounds || !matchings || !hits || !health || !damage) return fail("%s", "msb_env_shoot: null argument");
    if (subsample < 1 || res % subsample || s->n_model < 1) return fail("%s", "msb_env_shoot: subsample must divide res");
    if (s->n_envs == 0) return 0;
    shoot_kernel<<<s->n_envs, 128, (size_t)s->n_agents * sizeof(int), (cudaStream_t)cuda_stream>>>(
        indices, a->positions, bounds, matchings, hits, health, damage, s->n_envs, s->n_agents, res, subsample, s->n_model, clearance);
    g_launches++;
    return check(cudaGetLastError(), "shoot_kernel launch");
}

extern "C" int msb_env_respawn(const msb_scenery* s, const msb_agents* a, const uint8_t* reset, const float* spawn_positions,
                               const float* spawn_angles, int32_t n_spawns, uint32_t seed, uint32_t tick, const int32_t* choices,
                               void* cuda_stream) {
    if (!s || !a || !reset || !spawn_positions || !spawn_angles || n_spawns < 1) return fail("%s", "msb_env_respawn: bad argument");
    const int64_t total = (int64_t)s->n_envs * s->n_agents;
    if (total == 0) return 0;
    respawn_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(reset, spawn_positions, spawn_angles, *a, total, n_spawns,
                                                                                       seed, tick, const_cast<int32_t*>(choices));
    g_launches++;
    return check(cudaGetLastError(), "respawn_kernel launch");
}

extern "C" int msb_pack_obs(const float* rgb, const float* depth, const float* imu, int64_t n_envs, int32_t n_agents, int32_t ro,
                            void* rows, int64_t row_bytes, int32_t mode, int32_t imu_offset, void* cuda_stream) {
    if (!rgb || !depth || !imu || !rows || mode < 0 || mode > 2) return fail("%s", "msb_pack_obs: bad argument");
    if (n_envs == 0) return 0;
    const int n_rgb = n_agents * 3 * ro, n_d = n_agents * ro, n_imu = n_agents * 3;
    if (n_envs > 65535 * 1024ll) return fail("%s", "msb_pack_obs: too many envs for one launch");
    const int64_t out_align = mode == 0 ? 16 : (mode == 1 ? 8 : 4);          // four pixels of output
    const bool vec = n_rgb % 4 == 0 && n_d % 4 == 0 && row_bytes % out_align == 0 && ((uintptr_t)rows % out_align) == 0 &&
                     ((uintptr_t)rgb % 16) == 0 && ((uintptr_t)depth % 16) == 0;
    const int units = (n_rgb + n_d) / (vec ? 4 : 1) + n_imu;
    // (gridDim.y is capped at 65535: fold the envs of a larger batch into successive launches)
    for (int64_t n0 = 0; n0 < n_envs; n0 += 65535) {
        const int64_t nn = n_envs - n0 < 65535 ? n_envs - n0 : 65535;
        dim3 grid((unsigned)((units + 255) / 256), (unsigned)nn);
        unsigned char* out = reinterpret_cast<unsigned char*>(rows) + n0 * row_bytes;
        if (vec) pack_obs_kernel<true><<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(rgb + n0 * n_rgb, depth + n0 * n_d, imu + n0 * n_imu, out, n_rgb, n_d, n_imu, row_bytes, mode, imu_offset);
        else pack_obs_kernel<false><<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(rgb + n0 * n_rgb, depth + n0 * n_d, imu + n0 * n_imu, out, n_rgb, n_d, n_imu, row_bytes, mode, imu_offset);
        g_launches++;
    }
    return check(cudaGetLastError(), "pack_obs_kernel launch");
}
