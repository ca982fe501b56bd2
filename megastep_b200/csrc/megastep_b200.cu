// megastep_b200.cu — sm_100a kernels + C ABI (include/megastep_b200.h) for megastep's per-step hot path.
//
// Replaces, from scratch, the reference's megastep/src/kernels.cu:
//   physics()  = collision_kernel (:179-210) + ~15 ATen elementwise launches (:223-227)      -> ONE kernel
//   render()   = draw_kernel (:297-318) + raycast_kernel (:326-383) + shader_kernel (:407-450) -> ONE kernel
//   msb_step() = MomentumMovement (modules.py:106-118) + physics + render + RGB/Depth/IMU heads -> ONE kernel
//   bake()     = baking_kernel (:270-284)
//
// Layout / mapping:
//   * one CTA per environment; the env's static segments are staged once from the ragged-packed HBM array into
//     shared memory with a single 1-D bulk (TMA) copy + mbarrier, and shared by all of the env's agents;
//   * one warp per agent (x ray block): lanes are SEGMENTS while binning (each lane projects one segment onto the
//     agent's 1-D screen and gets the conservative interval of rays it can touch), then lanes are RAYS while
//     testing; a __ballot over "segment overlaps this 32-ray chunk" yields the candidates in ascending line order,
//     which preserves the reference's order-dependent nearest-hit rule exactly while skipping ~90% of the tests;
//   * per-(agent, segment) terms of the intersection are hoisted out of the per-ray work; the ray/line cosine and
//     its sqrt are computed for the winning line only (the reference computes them for every line);
//   * rays that hit another agent need the dynamic light at the hit point (I lights x W occluders); the warp does
//     that cooperatively (lanes = occluders, early exit on the first occluder) instead of one lane doing I*W tests;
//   * physics: the CTA's threads stride the env's segments for each agent, warp-shuffle min, fused integration.
//
// No tensor cores: nothing on this path is a dense contraction.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

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
    int32_t has_mv;
    int32_t ray_blocks;     // RB: warps per agent in the render stage
    int32_t seg_cap;        // float4 slots reserved for segments in shared memory (>= max_lines)
    int32_t wcap;           // view_kernel: slots for the env's padded run table (multiple of 16)
    float inv_fps;          // IEEE 1/fps (ATen's tensor/scalar == tensor*(1/scalar), kernels.cu:224,226)
    float mv_keep, mv_dv, mv_dw;   // 1-decay, accel/fps, ang_accel/fps evaluated in double like the Python does
    float inv_max_depth, inv_speed, inv_ang, inv_sub;   // reciprocals ATen would multiply by
    float bin_kappa, bin_rmid, bin_xclip;               // screen-space binning constants (culling only)
    unsigned long long* stats;  // optional diagnostics counters (may be null)
    int32_t debug_skip_dyn;     // timing experiments only: leave agent-hit rays unlit
    int32_t split_render;       // cast kernel writes the four scalar Render outputs, shade_kernel does the rest
    int32_t two_phase;          // render: bin every (agent, segment) once into shared memory, then one warp per ray chunk
    int32_t variant;            // bit 0: depth culling off; bit 1: software-pipelined candidate loop; bit 2: segment-major
                                // candidate stage (measured slower: 144 vs 134 us)
    // queue of pixel groups whose dynamic lighting is resolved by dyn_kernel (load-balanced second pass)
    int* dyn_ctrl;              // [0] entries reserved, [1] CTAs of dyn_kernel done
    unsigned char* dyn_entries; // null -> dynamic lights are resolved inline by the ray's own warp
    int32_t dyn_cap;            // entries that fit
    int32_t dyn_stride;         // bytes per entry = 16 + 32 * subsample
    int* dyn_cache;             // [N][A][32] last occluder of each light as seen from (around) each agent; a hint
};

enum { MODE_PHYSICS = 1, MODE_RENDER = 2, MODE_STEP = 3 };
#ifndef MSB_SHADE_ILP
#define MSB_SHADE_ILP 1       // chunks whose texel gathers are in flight together (2 spills at 64 registers, no gain)
#endif
#ifndef MSB_MIN_BLOCKS
#define MSB_MIN_BLOCKS 4      // 256 threads x 4 blocks -> at most 64 registers per thread
#endif
enum { ST_ANG = 0, ST_PX = 1, ST_PY = 2, ST_AV = 3, ST_VX = 4, ST_VY = 5, ST_STRIDE = 8 };
enum { STAT_TESTS = 0, STAT_GROUPS = 1, STAT_DYN_RAYS = 2, STAT_DYN_ITERS = 3, STAT_COLL = 4, STAT_REPLAYS = 5 };

struct Smem {
    float4* seg;        // [seg_cap] this env's segments {ax, ay, bx, by}
    float4* scratch;    // [nwarps][64] per-warp per-segment terms
    float* st_in;       // [A][8] start-of-step agent state
    float* st_out;      // [A][8] post-physics agent state
    int* xmin;          // [A] progress, as ordered int bits
    int* ncand;         // [A] physics: segments that survived the bounding-box cull
    unsigned short* cand;   // [A][seg_cap] their indices
    uint64_t* bar;
    long long* tstart;  // [seg_cap] render: this env's texel offsets (tex_starts) ...
    int* twidth;        // [seg_cap] ... and texel counts (tex_widths), so shading's first lookup is a shared-memory read
    float4* rec;        // two-phase render: [2][A][seg_cap] per-(agent, segment) records
};

__device__ __forceinline__ Smem carve(unsigned char* base, int seg_cap, int nwarps, int A, bool two_phase = false) {
    Smem m;
    m.seg = reinterpret_cast<float4*>(base);
    m.scratch = m.seg + seg_cap;
    m.st_in = reinterpret_cast<float*>(m.scratch + nwarps * 64);
    m.st_out = m.st_in + A * ST_STRIDE;
    m.xmin = reinterpret_cast<int*>(m.st_out + A * ST_STRIDE);
    m.ncand = m.xmin + A;
    m.cand = reinterpret_cast<unsigned short*>(m.ncand + A);
    uintptr_t p = reinterpret_cast<uintptr_t>(m.cand + (size_t)A * seg_cap);
    p = (p + 15) & ~uintptr_t(15);
    m.bar = reinterpret_cast<uint64_t*>(p);
    m.tstart = reinterpret_cast<long long*>(p + 16);
    m.twidth = reinterpret_cast<int*>(m.tstart + seg_cap);
    uintptr_t q = reinterpret_cast<uintptr_t>(m.twidth + seg_cap);
    q = (q + 15) & ~uintptr_t(15);
    m.rec = two_phase ? reinterpret_cast<float4*>(q) : nullptr;
    return m;
}

static size_t smem_bytes(int seg_cap, int nwarps, int A, bool two_phase = false) {
    size_t b = (size_t)seg_cap * 16 + (size_t)nwarps * 64 * 16 + (size_t)A * ST_STRIDE * 4 * 2 + (size_t)A * 8 +
               (size_t)A * seg_cap * 2;
    b = (b + 15) & ~size_t(15);
    b += 16 + (size_t)seg_cap * 12;
    b = (b + 15) & ~size_t(15);
    return b + (two_phase ? (size_t)2 * A * seg_cap * 16 : 0);
}

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

// One env's physics tick. st_in holds the start-of-step state of all A agents; results go to global memory and to
// st_out (for a following render stage).
template <bool BOXES>
__device__ __forceinline__ void physics_stage(const KArgs& k, const Smem& m, int n, int L) {
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31;
    const float rF = rcp(k.p.fps);
    const float r2 = fmul(k.p.agent_radius, 2.0020000934600830078f);
    const float r1 = fmul(k.p.agent_radius, 1.0010000467300415039f);
    const float r1sq = fmul(r1, r1);

    // One warp per agent, no block-level synchronisation.
    // pass 1: which static segments can possibly matter to this agent this tick? A segment farther from the agent
    // than rho = 1.05|v| + 2.2 r + 0.02 cannot trigger any branch of collision() with a result below 1 (see
    // DESIGN.md "physics cull"), so dropping it leaves progress bit-identical. Survivors are compacted (ballot +
    // prefix popcount) so that pass 2 runs the ~100-instruction test on dense lanes.
    const int cap = k.seg_cap;
    const int warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int a = warp; a < A; a += nwarps) {
        const float* me = m.st_in + a * ST_STRIDE;
        const float px = me[ST_PX], py = me[ST_PY], mx = me[ST_VX], my = me[ST_VY];
        const float vx = fmul(mx, rF), vy = fmul(my, rF);
        const float vlen = sqrt_(ffma(vx, vx, fmul(vy, vy)));
        // slow but moving agents: project()'s +1e-6 distorts distances -> test everything; exactly stationary ones can
        // only trigger the end-point branch (:163-168: every other branch needs s > 0), which the same radius covers
        const bool can_cull = vlen >= 1e-3f || (vx == 0.f && vy == 0.f);
        const float rho = 1.05f * vlen + 2.2f * r1 + 0.02f;
        float x = 1.f;
        // other agents (:193-200): start-of-step state, no sequential resolution
        for (int d1 = lane; d1 < A; d1 += 32) {
            if (d1 != a) {
                const float* o = m.st_in + d1 * ST_STRIDE;
                x = fminf(x, collide_agents(px, py, mx, my, o[ST_PX], o[ST_PY], o[ST_VX], o[ST_VY], rF, r2));
            }
        }
        const float u = fadd(vlen, 1e-6f), uu = fmul(u, u);
        if (BOXES) {
            // static lines (:203-205) from the occluder table (sorted copy of the env's static segments + a bounding
            // box per run): min over segments does not care about order. Lane b tests box b against the square the
            // agent can reach; only overlapping runs are read (straight from HBM/L2: no staging in this kernel).
            const int W = L - AF, run = k.s.occ_run, per = 32 / run;
            const int nb = (W + run - 1) / run;
            const float4* occ = reinterpret_cast<const float4*>(k.s.occ_lines) + __ldg(k.s.occ_starts + n);
            const float4* boxes = reinterpret_cast<const float4*>(k.s.occ_boxes) + __ldg(k.s.box_starts + n);
            const int slot = lane / run, within = lane - slot * run;
            for (int b0 = 0; b0 < nb; b0 += 32) {
                bool visit = false;
                if (b0 + lane < nb) {
                    const float4 bx = __ldg(boxes + b0 + lane);
                    visit = !(can_cull && (bx.x > px + rho || bx.z < px - rho || bx.y > py + rho || bx.w < py - rho));
                }
                unsigned runs = __ballot_sync(0xffffffffu, visit);
                while (runs) {
                    const unsigned rest = runs & (runs - 1);
                    const int n0 = __ffs(runs) - 1, n1 = rest ? __ffs(rest) - 1 : -1;
                    const int nth = (per == 1 || slot == 0) ? n0 : ((slot == 1) ? n1 : -1);
                    const int l = nth >= 0 ? run * (b0 + nth) + within : W;
                    if (l < W) {
                        const float4 s4 = __ldg(occ + l);
                        const bool outside = (fminf(s4.x, s4.z) > px + rho) || (fmaxf(s4.x, s4.z) < px - rho) ||
                                             (fminf(s4.y, s4.w) > py + rho) || (fmaxf(s4.y, s4.w) < py - rho);
                        if (!(can_cull && outside)) x = fminf(x, collide_line(px, py, vx, vy, vlen, u, uu, s4, r1, r1sq));
                    }
                    for (int d = 0; d < per && runs; d++) runs &= runs - 1;
                }
            }
        } else {
            unsigned short* mine = m.cand + a * cap;
            int nc = 0;
            for (int base = AF; base < L; base += 32) {
                const int l = base + lane;
                bool keep = false;
                if (l < L) {
                    const float4 s4 = m.seg[l];
                    const bool outside = (fminf(s4.x, s4.z) > px + rho) || (fmaxf(s4.x, s4.z) < px - rho) ||
                                         (fminf(s4.y, s4.w) > py + rho) || (fmaxf(s4.y, s4.w) < py - rho);
                    keep = !(can_cull && outside);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) mine[nc + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)l;
                nc += __popc(bal);
            }
            __syncwarp();
            // pass 2: exact tests on dense lanes; the agents' own model lines [0, AF) are never candidates
            for (int i = lane; i < nc; i += 32) {
                x = fminf(x, collide_line(px, py, vx, vy, vlen, u, uu, m.seg[mine[i]], r1, r1sq));
            }
        }
        x = warp_min(x);
        if (lane == 0) m.xmin[a] = __float_as_int(x);
    }
    __syncthreads();

    // integration (kernels.cu:223-227), one thread per agent, plain in-place stores (no storage swap)
    for (int a = tid; a < A; a += blockDim.x) {
        const float* me = m.st_in + a * ST_STRIDE;
        float* o = m.st_out + a * ST_STRIDE;
        const float x = __int_as_float(m.xmin[a]);
        const int64_t i = (int64_t)n * A + a;
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
    }
}

// ---------------------------------------------------------------------------------------------------------------
// render
// ---------------------------------------------------------------------------------------------------------------

// light_intensity() (kernels.cu:238-268) for a ray that hit an agent, evaluated by a whole warp with the first 32
// lights resident one per lane. Each lane remembers the static line that last occluded its light: consecutive
// agent-hit rays land centimetres apart, so one test per light (all lights in parallel) settles almost every
// occluded light; only the remaining lights are scanned against all static lines (lanes stride the lines, stop at
// the first occluder). Unoccluded lights are then accumulated in light order, exactly as the reference sums them.
// Which lines get tested varies; the occluded/unoccluded answer per light — hence the result — does not.
struct LaneLight { float x, y, i; int occ; };

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

// The same, over the occluder table, reading straight from HBM/L2: `occ` holds this env's W static segments sorted
// along a Morton curve, `boxes` the bounding box of each run of `run` of them. A run can only contain an occluder of
// the light ray I->C if its box, grown by a margin covering the worst-case rounding of intersect() (near-parallel
// lines: |UxV| >= 1e-3 bounds the amplification), overlaps the ray's box; all other runs are skipped. Lane b tests
// box b, a ballot gives the runs to visit, 32/run of them per warp iteration. Up to 32 lights, one per lane.
struct OccEnv { const float4* occ; const float4* boxes; int W, nb, run; float vmax, diam; };

template <bool STATS>
__device__ __forceinline__ float light_intensity_boxed(const OccEnv& oe, int I, float Cx, float Cy, int lane,
                                                       LaneLight& ll, unsigned& iters) {
    const int nres = I < 32 ? I : 32;
    bool ob = false;
    if (lane < nres && ll.occ >= 0 && ll.occ < oe.W) {
        const Hit h = intersect(ll.x, ll.y, fsub(Cx, ll.x), fsub(Cy, ll.y), __ldg(oe.occ + ll.occ));
        ob = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
    }
    const unsigned resident = nres == 32 ? 0xffffffffu : ((1u << nres) - 1u);
    unsigned todo = resident & ~__ballot_sync(0xffffffffu, ob);
    unsigned lit = 0;
    if (STATS) iters++;
    const int run = oe.run, per = 32 / run;          // segments per box; boxes scanned per warp iteration
    const int slot = lane / run, within = lane - slot * run;
    // this lane's boxes (box b0+lane of each 32-box round) do not depend on the light: load the first round once
    float4 bx0 = make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    if (todo && lane < oe.nb) bx0 = __ldg(oe.boxes + lane);
    while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        const float Ix = __shfl_sync(0xffffffffu, ll.x, i), Iy = __shfl_sync(0xffffffffu, ll.y, i);
        const float Ux = fsub(Cx, Ix), Uy = fsub(Cy, Iy);
        // conservative query box (see DESIGN.md "shadow cull"): rounding can move the computed crossing by at most
        // delta (a fraction of each segment's length) along either segment
        const float ulen = fmaxf(fabsf(Ux), fabsf(Uy));
        const float delta = 4e-4f * oe.vmax * (oe.diam + ulen);
        const float mg = delta * (ulen + oe.vmax) + 0.01f;
        const float qx0 = fminf(Ix, Cx) - mg, qx1 = fmaxf(Ix, Cx) + mg, qy0 = fminf(Iy, Cy) - mg, qy1 = fmaxf(Iy, Cy) + mg;
        int found = -1;
        for (int b0 = 0; b0 < oe.nb && found < 0; b0 += 32) {
            float4 bx = bx0;
            if (b0) bx = (b0 + lane < oe.nb) ? __ldg(oe.boxes + b0 + lane) : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
            const bool visit = !(bx.x > qx1 || bx.z < qx0 || bx.y > qy1 || bx.w < qy0);
            unsigned runs = __ballot_sync(0xffffffffu, visit);
            while (runs) {
                // lanes [slot*run, (slot+1)*run) take the slot-th box still to visit
                const unsigned rest = runs & (runs - 1);
                const int n0 = __ffs(runs) - 1, n1 = rest ? __ffs(rest) - 1 : -1;
                const int nth = (per == 1 || slot == 0) ? n0 : ((slot == 1) ? n1 : -1);
                const int l = nth >= 0 ? run * (b0 + nth) + within : oe.W;
                bool o = false;
                if (l < oe.W) {
                    const Hit h = intersect(Ix, Iy, Ux, Uy, __ldg(oe.occ + l));
                    o = (h.t > 0.f) && (h.t < 1.f) && (h.s > 0.f) && (h.s < .999f);
                }
                if (STATS) iters++;
                const unsigned bal = __ballot_sync(0xffffffffu, o);
                if (bal) { found = __shfl_sync(0xffffffffu, l, __ffs(bal) - 1); break; }
                for (int d = 0; d < per && runs; d++) runs &= runs - 1;    // drop the boxes just scanned
            }
        }
        if (found < 0) lit |= 1u << i;
        else if (lane == i) ll.occ = found;
    }
    float acc = 0.1f;   // AMBIENT (kernels.cu:9)
    while (lit) {
        const int i = __ffs(lit) - 1;
        lit &= lit - 1;
        const float Ix = __shfl_sync(0xffffffffu, ll.x, i), Iy = __shfl_sync(0xffffffffu, ll.y, i);
        const float Ii = __shfl_sync(0xffffffffu, ll.i, i);
        const float dx = fsub(Ix, Cx), dy = fsub(Iy, Cy);
        acc = ffma(fadd(Ii, Ii), rcp(fmaxf(ffma(dx, dx, fmul(dy, dy)), 1.f)), acc);   // LUMINANCE = 2 (:240)
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

// draw_kernel (kernels.cu:297-318): the agents' model lines, at their current poses, into shared and global memory.
__device__ __forceinline__ void draw_stage(const KArgs& k, const Smem& m, int n, int64_t g0) {
    const int A = k.s.n_agents, F = k.s.n_model;
    float* seg = reinterpret_cast<float*>(m.seg);
    for (int t = threadIdx.x; t < A * F * 2; t += blockDim.x) {
        const int e = t & 1, mm = (t >> 1) % F, a = (t >> 1) / F;
        const float* st = m.st_out + a * ST_STRIDE;
        float s, c;
        sincos_deg(st[ST_ANG], s, c);
        const float mx = __ldg(k.s.model + 4 * mm + 2 * e), my = __ldg(k.s.model + 4 * mm + 2 * e + 1);
        const float2 pt = make_float2(fadd(st[ST_PX], cross2(c, mx, s, my)), fadd(st[ST_PY], dot2(s, mx, c, my)));
        reinterpret_cast<float2*>(seg)[2 * (a * F + mm) + e] = pt;
        reinterpret_cast<float2*>(k.s.lines)[2 * (g0 + a * F + mm) + e] = pt;
    }
}

// shader_kernel (kernels.cu:407-450) for one 32-ray chunk of one agent's view, lane = ray, in two steps: shade_fetch
// (filter + texel/baked gathers) and shade_finish (dynamic light for agent hits — queued for dyn_kernel, or inline —
// the five Render outputs, and the fused Depth / RGB heads). `seg` = this env's segments (shared memory in the
// one-kernel render, HBM in shade_kernel).
struct ShadeCtx { int nlights; const float* lt; LaneLight ll; unsigned dyn_rays, dyn_iters; };

// What shading gathers for one ray: the two texels (and baked lights) around the hit, with the filter weights.
struct Texels { float lw, rw, tl0, tl1, tl2, tr0, tr1, tr2, bl, br; };

// filter() (kernels.cu:394-405) + the gathers of shader_kernel (:427-430, :438). `tw`/`ts` = texel count / offset per
// line of this env (shared memory in the one-kernel render, HBM in shade_kernel). Only issues loads; nothing here
// waits on them, so several chunks' gathers can be in flight before shade_finish() consumes the first.
__device__ __forceinline__ Texels shade_fetch(const KArgs& k, const int* __restrict__ tw, const long long* __restrict__ ts_,
                                              int AF, bool hitany, int l0, float locv) {
    Texels t;
    t.lw = t.rw = t.tl0 = t.tl1 = t.tl2 = t.tr0 = t.tr1 = t.tr2 = t.bl = t.br = 0.f;
    if (hitany) {
        const int w = tw[l0];
        const int64_t ts = ts_[l0];
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
        if (l0 >= AF) { t.bl = __ldg(k.s.baked + ts + fl); t.br = __ldg(k.s.baked + ts + fr); }
    }
    return t;
}

template <bool STATS>
__device__ __forceinline__ void shade_finish(const KArgs& k, const float4* __restrict__ seg, int n, int a, int L, int r,
                                             int lane, int l0, float locv, float dotv, float dist, const Texels& t,
                                             bool write_raw, ShadeCtx& sc_) {
    const int A = k.s.n_agents, AF = A * k.s.n_model, R = k.p.res;
    const int sub_ = k.has_obs ? k.obs.subsample : 1;
    const int nlights = sc_.nlights;
    const float* lt = sc_.lt;
    LaneLight& ll = sc_.ll;
    unsigned& dyn_rays = sc_.dyn_rays;
    unsigned& dyn_iters = sc_.dyn_iters;
    const bool live = r < R;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        const float lw = t.lw, rw = t.rw, tl0 = t.tl0, tl1 = t.tl1, tl2 = t.tl2, tr0 = t.tr0, tr1 = t.tr1, tr2 = t.tr2;
        float intensity = 0.f;
        float Cx = 0.f, Cy = 0.f;
        const bool hitany = live && (l0 >= 0);
        if (hitany) {
            if (l0 >= AF) {
                intensity = ffma(lw, t.bl, fmul(rw, t.br));                                             // :438
            } else {
                const float om = fsub(1.f, locv);                                                       // :435
                const float4 s4 = seg[l0];
                Cx = ffma(s4.x, om, fmul(locv, s4.z));
                Cy = ffma(s4.y, om, fmul(locv, s4.w));
            }
        }
        // (1 - dot^2) and the filtered texel, common to static and dynamic lighting (:442-445)
        const bool isdyn = hitany && (l0 < AF);
        float kk0 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
        if (hitany) {
            kk0 = ffma(-dotv, dotv, 1.f);
            b0 = ffma(lw, tl0, fmul(rw, tr0));
            b1 = ffma(lw, tl1, fmul(rw, tr1));
            b2 = ffma(lw, tl2, fmul(rw, tr2));
        }
        // dynamic lighting for rays that hit an agent's model (:434-436). Preferred: queue the pixel group for the
        // load-balanced second pass (dyn_kernel). Fallback (no workspace / queue full): this warp resolves them
        // one ray at a time.
        unsigned dm = __ballot_sync(0xffffffffu, isdyn);
        if (k.debug_skip_dyn) dm = 0;
        const int gl = lane & ~(sub_ - 1);                                    // first lane of my pixel group
        const unsigned subm = sub_ == 32 ? 0xffffffffu : ((1u << sub_) - 1u);
        const unsigned gmask = (dm >> gl) & subm;                             // my group's agent-hit pixels
        bool queued = false;
        if (dm && k.dyn_entries) {
            const unsigned leaders = __ballot_sync(0xffffffffu, gmask != 0 && lane == gl);
            const int cnt = __popc(leaders);
            int base = 0;
            if (lane == 0) base = atomicAdd(k.dyn_ctrl, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            queued = base + cnt <= k.dyn_cap;
            // which agent the group's first agent-hit pixel landed on: keys the persistent occluder cache
            const int tgt = __shfl_sync(0xffffffffu, l0, gl + (gmask ? __ffs(gmask) - 1 : 0)) / k.s.n_model;
            if (gmask) {
                const int slot = base + __popc(leaders & ((1u << gl) - 1u));
                if (slot < k.dyn_cap) {
                    unsigned char* e = k.dyn_entries + (size_t)slot * k.dyn_stride;
                    if (lane == gl) {
                        *reinterpret_cast<int4*>(e) = make_int4(n, a * R + (r - lane + gl), queued ? (int)gmask : 0, sub_ | (tgt << 8));
                    }
                    if (queued) {
                        float4* rec = reinterpret_cast<float4*>(e + 16) + 2 * (lane - gl);
                        rec[0] = make_float4(b0, b1, b2, kk0);
                        rec[1] = make_float4(Cx, Cy, intensity, isdyn ? 1.f : 0.f);
                    }
                }
            }
        }
        if (!queued) {
            while (dm) {
                const int j = __ffs(dm) - 1;
                dm &= dm - 1;
                const float cx = __shfl_sync(0xffffffffu, Cx, j), cy = __shfl_sync(0xffffffffu, Cy, j);
                const float v = light_intensity_cached<STATS>(seg, L, AF, nlights, lt, cx, cy, lane, ll, dyn_iters);
                if (lane == j) intensity = v;
                if (STATS) dyn_rays++;
            }
        }
        if (hitany) {
            const float kk = fmul(kk0, intensity);
            s0 = fmul(kk, b0);
            s1 = fmul(kk, b1);
            s2 = fmul(kk, b2);
        }
        const bool deferred = queued && gmask != 0;      // dyn_kernel writes this group's screen / rgb
        if (live) {
            const int64_t o = ((int64_t)n * A + a) * R + r;
            if (write_raw) {
                if (k.out.indices) k.out.indices[o] = l0;
                if (k.out.locations) k.out.locations[o] = locv;
                if (k.out.dots) k.out.dots[o] = dotv;
                if (k.out.distances) k.out.distances[o] = dist;
            }
            if (k.out.screen && !(queued && isdyn)) { float* sc = k.out.screen + 3 * o; sc[0] = s0; sc[1] = s1; sc[2] = s2; }
        }
        // fused observation heads: Depth (modules.py:181-183) and RGB (:222-223), mean over `subsample` pixels
        if (k.has_obs) {
            float d = 0.f;
            if (live) {
                const float z = __fmul_rn(__fsub_rn(dist, k.p.agent_radius), k.inv_max_depth);
                d = __fsub_rn(1.f, fminf(fmaxf(z, 0.f), 1.f));
            }
            float v0 = s0, v1 = s1, v2 = s2, v3 = d;
            for (int o = 1; o < sub_; o <<= 1) {
                v0 = __fadd_rn(v0, __shfl_xor_sync(0xffffffffu, v0, o));
                v1 = __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                v2 = __fadd_rn(v2, __shfl_xor_sync(0xffffffffu, v2, o));
                v3 = __fadd_rn(v3, __shfl_xor_sync(0xffffffffu, v3, o));
            }
            if (live && lane == gl) {
                const int Ro = R / sub_, ro = r / sub_;
                const float inv = k.inv_sub;
                const int64_t ag = (int64_t)n * A + a;
                if (k.obs.rgb && !deferred) {
                    float* q = k.obs.rgb + ag * 3 * Ro + ro;
                    q[0] = __fmul_rn(v0, inv); q[Ro] = __fmul_rn(v1, inv); q[2 * Ro] = __fmul_rn(v2, inv);
                }
                if (k.obs.depth) k.obs.depth[ag * Ro + ro] = __fmul_rn(v3, inv);
            }
        }
}

// Per-(agent, segment) work shared by every ray of the agent: the exact ray-independent terms of intersect()
// (kernels.cu:83-85: V, PQ, cross(PQ, V)) and a CONSERVATIVE screen-space summary used only to skip work — the
// interval [rlo, rhi] of rays the segment can touch and a lower bound smin of the hit parameter s over it.
struct SegBin { float4 q0; float snum; float smin; int rlo, rhi; };

__device__ __forceinline__ SegBin bin_segment(const KArgs& k, float4 s4, float px, float py, float cs, float sn,
                                              float lo_ray, float hi_ray) {
    SegBin b;
    const float Vx = fsub(s4.z, s4.x), Vy = fsub(s4.w, s4.y);
    const float PQx = fsub(s4.x, px), PQy = fsub(s4.y, py);
    b.q0 = make_float4(Vx, Vy, PQx, PQy);
    b.snum = cross2(Vy, PQx, Vx, PQy);
    b.rlo = 1; b.rhi = 0; b.smin = CUDART_INF_F;
    // approximate camera-space endpoints: x' forward, y' left; screen coordinate = y'/x'
    const float kappa = k.bin_kappa;     // rays per unit of screen coordinate, R / (2 tan(fov/2))
    const float rmid = k.bin_rmid;       // (R - 1) / 2
    const float xclip = k.bin_xclip;     // well inside every ray's near plane
    const float delta = 0.05f;
    const float bxr = s4.z - px, byr = s4.w - py;
    float xa = PQx * cs + PQy * sn, ya = PQy * cs - PQx * sn;
    float xb = bxr * cs + byr * sn, yb = byr * cs - bxr * sn;
    const bool behind = (xa < xclip) && (xb < xclip);
    if (!behind) {
        if (xa < xclip) { const float tt = __fdividef(xclip - xa, xb - xa); ya = ya + tt * (yb - ya); xa = xclip; }
        if (xb < xclip) { const float tt = __fdividef(xclip - xb, xa - xb); yb = yb + tt * (ya - yb); xb = xclip; }
        const float sa = __fdividef(ya, xa), sb = __fdividef(yb, xb);
        const float rf_first = rmid - fmaxf(sa, sb) * kappa - delta;
        const float rf_last = rmid - fminf(sa, sb) * kappa + delta;
        // fmaxf/fminf return the non-NaN operand, so a NaN leaves the full range and the exact test still runs
        b.rlo = (int)fminf(fmaxf(ceilf(rf_first), lo_ray), hi_ray + 1.f);
        b.rhi = (int)fmaxf(fminf(floorf(rf_last), hi_ray), lo_ray - 1.f);
        b.smin = fminf(xa, xb) - 1e-3f - 1e-4f * fmaxf(fabsf(xa), fabsf(xb));
    }
    return b;
}

template <int NCH, bool STATS, bool SPLIT>
__device__ __forceinline__ void render_agent(const KArgs& k, const Smem& m, int n, int64_t g0, int L, int a, int rb,
                                             float4* __restrict__ scr, int lane) {
    // two-phase mode: the per-(agent, segment) records were written to shared memory by the binning phase
    const bool pre = m.rec != nullptr;
    const float4* rec0 = pre ? m.rec + (size_t)a * k.seg_cap : nullptr;
    const float4* rec1 = pre ? m.rec + (size_t)(k.s.n_agents + a) * k.seg_cap : nullptr;
    const int A = k.s.n_agents, AF = A * k.s.n_model, R = k.p.res;
    const float* st = m.st_out + a * ST_STRIDE;
    const float px = st[ST_PX], py = st[ST_PY];
    float sn, cs;
    sincos_deg(st[ST_ANG], sn, cs);

    // ---- rays (kernels.cu:341-344, ray_y :234-236). Lane = ray within each of this warp's NCH 32-ray chunks.
    const float Rf = (float)R;
    const float rcpR = rcp(Rf);
    const int r0 = rb * (32 * NCH);
    float rux[NCH], ruy[NCH], rlen[NCH], nearp[NCH], best[NCH], bestm[NCH], loc[NCH];
    int idx[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int r = r0 + 32 * c + lane;
        const float y = fmul(fmul(fadd(fsub(Rf, (float)(unsigned)(2 * r)), -1.f), k.p.half_screen), rcpR);
        rux[c] = ffma(sn, -y, cs);
        ruy[c] = ffma(cs, y, sn);
        rlen[c] = sqrt_(ffma(rux[c], rux[c], fmul(ruy[c], ruy[c])));
        nearp[c] = fmul(rcp(rlen[c]), k.p.agent_radius);
        best[c] = CUDART_INF_F;
        bestm[c] = CUDART_INF_F;   // best + -1e-4f, kept alongside (inf - 1e-4 = inf)
        loc[c] = __int_as_float(0x7fffffff);
        idx[c] = -1;
    }

    const float lo_chunk = (float)r0, hi_chunk = (float)(r0 + 32 * NCH - 1);
    unsigned tests = 0, groups = 0;

    for (int gbase = 0; gbase < L; gbase += 32) {
        const int l = gbase + lane;
        int rlo = 1, rhi = 0;
        float smin = CUDART_INF_F;          // conservative lower bound of s (= forward distance) over the segment
        const float4* q0p = scr;            // candidate j's terms: q0p[j] = {V, PQ}, q1p[j].x = cross(PQ, V)
        const float4* q1p = scr + 32;
        if (pre) {
            q0p = rec0 + gbase;
            q1p = rec1 + gbase;
            if (l < L) {
                const float4 r1 = q1p[lane];
                smin = r1.y; rlo = __float_as_int(r1.z); rhi = __float_as_int(r1.w);
            }
        } else {
            if (l < L) {
                const SegBin b = bin_segment(k, m.seg[l], px, py, cs, sn, lo_chunk, hi_chunk);
                scr[lane] = b.q0;
                scr[32 + lane].x = b.snum;
                rlo = b.rlo; rhi = b.rhi; smin = b.smin;
            }
            __syncwarp();
        }
        if (!(k.variant & 4)) {
    #pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int c_lo = r0 + 32 * c, c_hi = c_lo + 31;
                // depth cull (exact): a hit on this segment has s >= smin; if that is beyond the current hit of EVERY ray
                // of the chunk it cannot satisfy s < best - 1e-4 for any of them
                float cmax = CUDART_INF_F;
                if (!(k.variant & 1)) {
                    // best >= 0 (or +inf), so its bit pattern orders like an unsigned integer: one REDUX instead of a
                    // shuffle tree
                    cmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best[c])));
                }
                unsigned mask = __ballot_sync(0xffffffffu, (rlo <= c_hi) && (rhi >= c_lo) && (rlo <= rhi) && !(smin > cmax));
                if (k.variant & 2) {
                    if (mask) {
                        int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        float4 q = q0p[j];
                        float snum = q1p[j].x;
                        while (true) {
                            // fetch the next candidate's terms before the arithmetic of this one (hides the LDS latency)
                            const bool more = mask != 0;
                            const int jn = more ? __ffs(mask) - 1 : j;
                            mask &= mask - 1;
                            const float4 qn = q0p[jn];
                            const float snn = q1p[jn].x;
                            const float UxV = cross2(rux[c], q.y, ruy[c], q.x);
                            const float rc = rcp(UxV);
                            const float hs_ = fmul(snum, rc);
                            const float ht_ = fmul(cross2(ruy[c], q.z, rux[c], q.w), rc);
                            const bool take = !(fabsf(UxV) < PARALLEL_EPS) && (ht_ >= 0.f) && (ht_ <= 1.f) && (nearp[c] < hs_) && (hs_ < bestm[c]);
                            if (take) { best[c] = hs_; bestm[c] = fadd(hs_, -1.e-4f); loc[c] = ht_; idx[c] = gbase + j; }
                            if (STATS) tests++;
                            if (!more) break;
                            j = jn; q = qn; snum = snn;
                        }
                    }
                } else {
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float4 q = q0p[j];
                        const float snum = q1p[j].x;
                        // raycast_kernel inner loop (kernels.cu:353-376), branch-free. A near-parallel line
                        // (|UxV| < 1e-3: s = t = inf in the reference) can never be accepted, so its s/t need no forcing.
                        const float UxV = cross2(rux[c], q.y, ruy[c], q.x);
                        const float rc = rcp(UxV);
                        const float hs_ = fmul(snum, rc);
                        const float ht_ = fmul(cross2(ruy[c], q.z, rux[c], q.w), rc);
                        const bool take = !(fabsf(UxV) < PARALLEL_EPS) && (ht_ >= 0.f) && (ht_ <= 1.f) && (nearp[c] < hs_) && (hs_ < bestm[c]);
                        if (take) { best[c] = hs_; bestm[c] = fadd(hs_, -1.e-4f); loc[c] = ht_; idx[c] = gbase + j; }
                        if (STATS) tests++;
                    }
                }
            }
        } else {
            // Segment-major candidate stage. cm = the chunks of this warp's ray block that my segment can still matter
            // to: its ray interval overlaps the chunk, and (depth cull, exact) its nearest point is not behind the
            // current hit of EVERY ray of the chunk — a hit on it has s >= smin, so it could not satisfy
            // s < best - 1e-4 for any of them. One ballot finds the segments to visit; each is loaded once and tested
            // against all its chunks. Rays still meet their candidates in ascending line order.
            unsigned cm = 0;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int c_lo = r0 + 32 * c, c_hi = c_lo + 31;
                // best >= 0 (or +inf): its bit pattern orders like an unsigned integer, one REDUX gives the chunk max
                const float cmax = (k.variant & 1) ? CUDART_INF_F
                                                   : __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best[c])));
                if ((rlo <= c_hi) && (rhi >= c_lo) && (rlo <= rhi) && !(smin > cmax)) cm |= 1u << c;
            }
            unsigned mask = __ballot_sync(0xffffffffu, cm != 0);
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1;
                const unsigned cj = __shfl_sync(0xffffffffu, cm, j);
                const float4 q = q0p[j];
                const float snum = q1p[j].x;
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    if (cj & (1u << c)) {
                        // raycast_kernel inner loop (kernels.cu:353-376), branch-free per lane. A near-parallel line
                        // (|UxV| < 1e-3: s = t = inf in the reference) can never be accepted: its s/t need no forcing.
                        const float UxV = cross2(rux[c], q.y, ruy[c], q.x);
                        const float rc = rcp(UxV);
                        const float hs_ = fmul(snum, rc);
                        const float ht_ = fmul(cross2(ruy[c], q.z, rux[c], q.w), rc);
                        const bool take = !(fabsf(UxV) < PARALLEL_EPS) && (ht_ >= 0.f) && (ht_ <= 1.f) && (nearp[c] < hs_) && (hs_ < bestm[c]);
                        if (take) { best[c] = hs_; bestm[c] = fadd(hs_, -1.e-4f); loc[c] = ht_; idx[c] = gbase + j; }
                        if (STATS) tests++;
                    }
                }
            }
        }
        if (STATS) groups++;
        if (!pre) __syncwarp();
    }

    // ---- per chunk: the winner's ray . line cosine (kernels.cu:362-364, winner only), then either hand the hit to
    // shade_kernel through the four scalar Render outputs (split render) or shade right here
    float dots_[NCH];
    ShadeCtx sc_;
    sc_.nlights = __ldg(k.s.light_widths + n);
    sc_.lt = k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n);
    sc_.ll.occ = -1;
    sc_.dyn_rays = sc_.dyn_iters = 0;
    if (!SPLIT && lane < sc_.nlights) { sc_.ll.x = __ldg(sc_.lt + 3 * lane); sc_.ll.y = __ldg(sc_.lt + 3 * lane + 1); sc_.ll.i = __ldg(sc_.lt + 3 * lane + 2); }
    else { sc_.ll.x = 0.f; sc_.ll.y = 0.f; sc_.ll.i = 0.f; }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        const int r = r0 + 32 * c + lane;
        const int l0 = idx[c];
        float dotv = __int_as_float(0x7fffffff);
        if (l0 >= 0) {
            const float4 s4 = m.seg[l0];
            const float Vx = fsub(s4.z, s4.x), Vy = fsub(s4.w, s4.y);
            dotv = fmul(dot2(rux[c], Vx, ruy[c], Vy), rcp(ffma(rlen[c], sqrt_(ffma(Vx, Vx, fmul(Vy, Vy))), 1.e-6f)));
        }
        const float dist = fmul(rlen[c], best[c]);
        if (SPLIT) {
            if (r < R) {
                const int64_t o = ((int64_t)n * A + a) * R + r;
                k.out.indices[o] = l0;
                k.out.locations[o] = loc[c];
                k.out.dots[o] = dotv;
                k.out.distances[o] = dist;
            }
        } else {
            dots_[c] = dotv;
        }
    }
    if (!SPLIT) {
        // gathers of two chunks in flight at a time, then their arithmetic, stores and queueing
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += MSB_SHADE_ILP) {
            Texels tx[MSB_SHADE_ILP];
#pragma unroll
            for (int u = 0; u < MSB_SHADE_ILP; u++) {
                const int c = c0 + u;
                if (c < NCH) tx[u] = shade_fetch(k, m.twidth, m.tstart, AF, (r0 + 32 * c + lane < R) && idx[c] >= 0, idx[c], loc[c]);
            }
#pragma unroll
            for (int u = 0; u < MSB_SHADE_ILP; u++) {
                const int c = c0 + u;
                if (c < NCH)
                    shade_finish<STATS>(k, m.seg, n, a, L, r0 + 32 * c + lane, lane, idx[c], loc[c], dots_[c],
                                        fmul(rlen[c], best[c]), tx[u], true, sc_);
            }
        }
    }
    const unsigned dyn_rays = sc_.dyn_rays, dyn_iters = sc_.dyn_iters;
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_TESTS, (unsigned long long)tests);
        atomicAdd(k.stats + STAT_GROUPS, (unsigned long long)groups);
        atomicAdd(k.stats + STAT_DYN_RAYS, (unsigned long long)dyn_rays);
        atomicAdd(k.stats + STAT_DYN_ITERS, (unsigned long long)dyn_iters);
    }
}

// IMU head (modules.py:263-270): {angvelocity/ang_scale, to_local_frame(angles, velocity)/speed_scale}
__device__ __forceinline__ void imu_stage(const KArgs& k, const Smem& m, int n) {
    const int A = k.s.n_agents;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
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

// ---------------------------------------------------------------------------------------------------------------
// the per-env kernel: any of physics / render / both
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int NCH, bool STATS, bool SPLIT>
__global__ void __launch_bounds__(256, MSB_MIN_BLOCKS) env_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const Smem m = carve(smem_raw, k.seg_cap, nwarps, A, (MODE & MODE_RENDER) && k.two_phase);

    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    const int W = L - AF;

    // physics on its own reads the few segments it needs from the occluder table and skips the staging altogether
    const bool boxes = (MODE == MODE_PHYSICS) && k.s.occ_lines != nullptr;
    // stage this env's static segments: one bulk (TMA) copy, ragged-packed HBM -> shared memory
    if (tid == 0 && !boxes) {
        mbar_init(m.bar, 1);
        if (W > 0) {
            mbar_expect_tx(m.bar, (uint32_t)W * 16u);
            bulk_g2s(m.seg + AF, k.s.lines + 4 * (g0 + AF), (uint32_t)W * 16u, m.bar);
        }
    }
    // agent state -> shared memory (and the MomentumMovement update when fused, modules.py:106-118)
    for (int a = tid; a < A; a += blockDim.x) {
        const int64_t i = (int64_t)n * A + a;
        float ang = k.a.angles[i], av = k.a.angvelocity[i];
        float2 pos = reinterpret_cast<const float2*>(k.a.positions)[i];
        float2 vel = reinterpret_cast<const float2*>(k.a.velocity)[i];
        if ((MODE & MODE_PHYSICS) && k.has_mv) {
            const int act = k.mv.actions[i];
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
        float* st = ((MODE & MODE_PHYSICS) ? m.st_in : m.st_out) + a * ST_STRIDE;
        st[ST_ANG] = ang; st[ST_PX] = pos.x; st[ST_PY] = pos.y; st[ST_AV] = av; st[ST_VX] = vel.x; st[ST_VY] = vel.y;
        m.xmin[a] = __float_as_int(1.f);
        m.ncand[a] = 0;
    }
    if (MODE & MODE_RENDER) {
        for (int l = tid; l < L; l += blockDim.x) {
            m.twidth[l] = __ldg(k.s.tex_widths + g0 + l);
            m.tstart[l] = __ldg(reinterpret_cast<const long long*>(k.s.tex_starts) + g0 + l);
        }
    }
    __syncthreads();
    if (W > 0 && !boxes) mbar_wait(m.bar, 0);

    if (MODE & MODE_PHYSICS) {
        if (boxes) physics_stage<true>(k, m, n, L);
        else physics_stage<false>(k, m, n, L);
        if (MODE & MODE_RENDER) __syncthreads();
    }
    if (MODE & MODE_RENDER) {
        draw_stage(k, m, n, g0);
        __syncthreads();
        const int RB = k.ray_blocks;
        if (m.rec) {
            // phase 1 of two-phase render: every (agent, segment) binned once, by whichever warp gets to it
            const int ngroups = (L + 31) >> 5;
            const float hi_ray = (float)(k.p.res - 1);
            for (int item = warp; item < A * ngroups; item += nwarps) {
                const int a = item / ngroups, l = 32 * (item - a * ngroups) + lane;
                const float* st = m.st_out + a * ST_STRIDE;
                float sn, cs;
                sincos_deg(st[ST_ANG], sn, cs);
                if (l < L) {
                    const SegBin b = bin_segment(k, m.seg[l], st[ST_PX], st[ST_PY], cs, sn, 0.f, hi_ray);
                    m.rec[(size_t)a * k.seg_cap + l] = b.q0;
                    m.rec[(size_t)(A + a) * k.seg_cap + l] = make_float4(b.snum, b.smin, __int_as_float(b.rlo), __int_as_float(b.rhi));
                }
            }
            __syncthreads();
        }
        for (int w = warp; w < A * RB; w += nwarps) {
            render_agent<NCH, STATS, SPLIT>(k, m, n, g0, L, w / RB, w % RB, m.scratch + warp * 64, lane);
        }
        if (k.has_obs && k.obs.imu) imu_stage(k, m, n);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shade_kernel: second half of the split render. One warp per (agent, 32-ray chunk), lane = ray; reads the hit the
// cast kernel left in the Render outputs and does everything that is memory-latency bound (texel / baked-light
// gathers, the queue for agent hits, screen, Depth/RGB heads) at full occupancy, with no shared memory.
// ---------------------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(256) shade_kernel(const __grid_constant__ KArgs k) {
    const int lane = threadIdx.x & 31;
    const int A = k.s.n_agents, R = k.p.res;
    const int chunks = (R + 31) >> 5;
    const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = (int64_t)k.s.n_envs * A * chunks;
    if (wg >= total) return;
    const int64_t ag = wg / chunks;
    const int c = (int)(wg - ag * chunks);
    const int n = (int)(ag / A), a = (int)(ag - (int64_t)n * A);
    const int r = 32 * c + lane;
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    const float4* seg = reinterpret_cast<const float4*>(k.s.lines) + g0;
    int l0 = -1;
    float locv = __int_as_float(0x7fffffff), dotv = __int_as_float(0x7fffffff), dist = CUDART_INF_F;
    if (r < R) {
        const int64_t o = ag * R + r;
        l0 = k.out.indices[o];
        locv = k.out.locations[o];
        dotv = k.out.dots[o];
        dist = k.out.distances[o];
    }
    ShadeCtx sc_;
    sc_.nlights = __ldg(k.s.light_widths + n);
    sc_.lt = k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n);
    sc_.ll.occ = -1;
    sc_.dyn_rays = sc_.dyn_iters = 0;
    sc_.ll.x = sc_.ll.y = sc_.ll.i = 0.f;
    if (lane < sc_.nlights) {                        // for the inline fallback (no workspace, or the queue is full)
        sc_.ll.x = __ldg(sc_.lt + 3 * lane); sc_.ll.y = __ldg(sc_.lt + 3 * lane + 1); sc_.ll.i = __ldg(sc_.lt + 3 * lane + 2);
    }
    const int AFk = A * k.s.n_model;
    const Texels tx = shade_fetch(k, k.s.tex_widths + g0, reinterpret_cast<const long long*>(k.s.tex_starts) + g0, AFk,
                                  (r < R) && l0 >= 0, l0, locv);
    shade_finish<STATS>(k, seg, n, a, L, r, lane, l0, locv, dotv, dist, tx, false, sc_);
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_DYN_RAYS, (unsigned long long)sc_.dyn_rays);
        atomicAdd(k.stats + STAT_DYN_ITERS, (unsigned long long)sc_.dyn_iters);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// view_kernel: render over the env's spatial table (Morton-sorted static segments in runs of 16 with bounding boxes).
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
enum { VRUN = 16 };

struct VSmem {
    float4* seg;            // [AF + wcap]: [0, AF) the agents' model lines at their current poses; then the sorted static rows
    float4* boxes;          // [wcap / 16]
    int4* rec;              // [wcap] per sorted row: {texel offset lo, hi, texel count, line id}
    float4* scr;            // [nwarps][128] per warp: 64 candidate records while casting, then the chunk results
    float* st_in;           // [A][8]
    float* st_out;          // [A][8]
    int* xmin;              // [A]
    uint64_t* bar;
};

__device__ __forceinline__ VSmem vcarve(unsigned char* base, int wcap, int nwarps, int A, int AF) {
    VSmem m;
    m.seg = reinterpret_cast<float4*>(base);
    m.boxes = m.seg + AF + wcap;
    m.rec = reinterpret_cast<int4*>(m.boxes + wcap / VRUN);
    m.scr = reinterpret_cast<float4*>(m.rec + wcap);
    m.st_in = reinterpret_cast<float*>(m.scr + nwarps * 128);
    m.st_out = m.st_in + A * ST_STRIDE;
    m.xmin = reinterpret_cast<int*>(m.st_out + A * ST_STRIDE);
    uintptr_t p = reinterpret_cast<uintptr_t>(m.xmin + A);
    p = (p + 15) & ~uintptr_t(15);
    m.bar = reinterpret_cast<uint64_t*>(p);
    return m;
}

static size_t vsmem_bytes(int wcap, int nwarps, int A, int AF) {
    size_t b = (size_t)(AF + wcap) * 16 + (size_t)(wcap / VRUN) * 16 + (size_t)nwarps * 128 * 16 + (size_t)wcap * 16 +
               (size_t)A * ST_STRIDE * 4 * 2 + (size_t)A * 4;
    b = (b + 15) & ~size_t(15);
    return b + 16;
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
            const int4 rc = m.rec[row - AF];
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

__device__ __forceinline__ void shade_chunk(const KArgs& k, const float4* __restrict__ seg, int n, int a, int AF, int Lrows,
                                            int r, int lane, const ShadeIn& in, const Texels& t) {
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
    // rays that hit an agent's model need the light at the hit point (:434-436): queue the pixel group for dyn_kernel
    unsigned dm = __ballot_sync(0xffffffffu, isdyn);
    if (k.debug_skip_dyn) dm = 0;
    const int gl = lane & ~(sub_ - 1);                                    // first lane of my pixel group
    bool queued = false, deferred = false;
    if (dm) {
        const unsigned subm = sub_ == 32 ? 0xffffffffu : ((1u << sub_) - 1u);
        const unsigned gmask = (dm >> gl) & subm;                         // my group's agent-hit pixels
        if (k.dyn_entries) {
            const unsigned leaders = __ballot_sync(0xffffffffu, gmask != 0 && lane == gl);
            const int cnt = __popc(leaders);
            int base = 0;
            if (lane == 0) base = atomicAdd(k.dyn_ctrl, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            queued = base + cnt <= k.dyn_cap;
            // which agent the group's first agent-hit pixel landed on: keys the persistent occluder cache
            const int tgt = __shfl_sync(0xffffffffu, l0, gl + (gmask ? __ffs(gmask) - 1 : 0)) / k.s.n_model;
            if (gmask) {
                const int slot = base + __popc(leaders & ((1u << gl) - 1u));
                if (slot < k.dyn_cap) {
                    unsigned char* e = k.dyn_entries + (size_t)slot * k.dyn_stride;
                    if (lane == gl) *reinterpret_cast<int4*>(e) = make_int4(n, a * R + (r - lane + gl), queued ? (int)gmask : 0, sub_ | (tgt << 8));
                    if (queued) {
                        float4* rec = reinterpret_cast<float4*>(e + 16) + 2 * (lane - gl);
                        rec[0] = make_float4(b0, b1, b2, kk0);
                        rec[1] = make_float4(Cx, Cy, intensity, isdyn ? 1.f : 0.f);
                    }
                }
            }
        }
        if (!queued) intensity = dyn_inline(seg, Lrows, AF, __ldg(k.s.light_widths + n), k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n), dm, Cx, Cy, intensity);
        deferred = queued && gmask != 0;                                  // dyn_kernel writes this group's screen / rgb
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (hitany) {
        const float kk = fmul(kk0, intensity);
        s0 = fmul(kk, b0);
        s1 = fmul(kk, b1);
        s2 = fmul(kk, b2);
    }
    const int64_t ag = (int64_t)n * A + a;
    if (live) {
        const int64_t o = ag * R + r;
        if (k.out.indices) k.out.indices[o] = l0;
        if (k.out.locations) k.out.locations[o] = locv;
        if (k.out.dots) k.out.dots[o] = dotv;
        if (k.out.distances) k.out.distances[o] = dist;
        if (k.out.screen && !(queued && isdyn)) { float* sc = k.out.screen + 3 * o; sc[0] = s0; sc[1] = s1; sc[2] = s2; }
    }
    // fused observation heads: Depth (modules.py:181-183) and RGB (:222-223), mean over `subsample` pixels
    if (k.has_obs) {
        float d = 0.f;
        if (live) {
            const float z = __fmul_rn(__fsub_rn(dist, k.p.agent_radius), k.inv_max_depth);
            d = __fsub_rn(1.f, fminf(fmaxf(z, 0.f), 1.f));
        }
        float v0 = s0, v1 = s1, v2 = s2, v3 = d;
        for (int o = 1; o < sub_; o <<= 1) {
            v0 = __fadd_rn(v0, __shfl_xor_sync(0xffffffffu, v0, o));
            v1 = __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, o));
            v2 = __fadd_rn(v2, __shfl_xor_sync(0xffffffffu, v2, o));
            v3 = __fadd_rn(v3, __shfl_xor_sync(0xffffffffu, v3, o));
        }
        if (live && lane == gl) {
            const int Ro = R / sub_, ro = r / sub_;
            const float inv = k.inv_sub;
            if (k.obs.rgb && !deferred) {
                float* q = k.obs.rgb + ag * 3 * Ro + ro;
                q[0] = __fmul_rn(v0, inv); q[Ro] = __fmul_rn(v1, inv); q[2 * Ro] = __fmul_rn(v2, inv);
            }
            if (k.obs.depth) k.obs.depth[ag * Ro + ro] = __fmul_rn(v3, inv);
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
    sincos_deg(st[ST_ANG], v.sn, v.cs);
    v.xclip = k.bin_xclip;
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
            packed = (row << 16) | (row < AF ? row : m.rec[row - AF].w);
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
            if (c0 + u < NCH) shade_chunk(k, m.seg, n, a, AF, AF + W, r0 + 32 * (c0 + u) + lane, lane, in[u], tx[u]);
        }
    }
    __syncwarp();
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_TESTS, (unsigned long long)tests);
        atomicAdd(k.stats + STAT_GROUPS, (unsigned long long)groups);
        atomicAdd(k.stats + STAT_REPLAYS, (unsigned long long)replays);
    }
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
__device__ __forceinline__ void physics_integrate(const KArgs& k, const float* st_in, float* st_out, int n, int a, float x) {
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

template <int NCH, bool PHYS, bool STATS>
__global__ void __launch_bounds__(256, MSB_MIN_BLOCKS) view_kernel(const __grid_constant__ KArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int A = k.s.n_agents, AF = A * k.s.n_model;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const VSmem m = vcarve(smem_raw, k.wcap, nwarps, A, AF);
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    const int W = L - AF;
    const int nb = (W + VRUN - 1) / VRUN;
    // stage this env's table: three bulk (TMA) copies, ragged-packed HBM -> shared memory, one mbarrier
    if (tid == 0) {
        mbar_init(m.bar, 1);
        if (nb > 0) {
            const int64_t b0 = __ldg(k.s.box_starts + n);
            mbar_expect_tx(m.bar, (uint32_t)nb * (VRUN * 32u + 16u));
            bulk_g2s(m.seg + AF, k.s.occ_lines + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u, m.bar);
            bulk_g2s(m.rec, k.s.occ_rec + 4 * VRUN * b0, (uint32_t)nb * VRUN * 16u, m.bar);
            bulk_g2s(m.boxes, k.s.occ_boxes + 4 * b0, (uint32_t)nb * 16u, m.bar);
        }
    }
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
            if (lane == 0) physics_integrate(k, m.st_in, m.st_out, n, a, x);
        }
        __syncthreads();
    }
    // draw_kernel (kernels.cu:297-318): the agents' model lines, at their current poses, into shared and global memory
    {
        const int F = k.s.n_model;
        for (int a = warp; a < A; a += nwarps) {
            const float* st = m.st_out + a * ST_STRIDE;
            float sn, cs;
            sincos_deg(st[ST_ANG], sn, cs);
            for (int t = lane; t < 2 * F; t += 32) {            // t = 2 * model line + endpoint
                const float2 mp = __ldg(reinterpret_cast<const float2*>(k.s.model) + t);
                const float2 pt = make_float2(fadd(st[ST_PX], cross2(cs, mp.x, sn, mp.y)), fadd(st[ST_PY], dot2(sn, mp.x, cs, mp.y)));
                reinterpret_cast<float2*>(m.seg)[2 * a * F + t] = pt;
                reinterpret_cast<float2*>(k.s.lines)[2 * (g0 + a * F) + t] = pt;
            }
        }
    }
    __syncthreads();
    const int RB = k.ray_blocks;
    for (int w = warp; w < A * RB; w += nwarps) {
        view_agent<NCH, STATS>(k, m, n, g0, L, W, nb, w / RB, w % RB, m.scr + warp * 128, lane);
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
}

// ---------------------------------------------------------------------------------------------------------------
// dyn_kernel: the load-balanced second pass over pixel groups that contain agent-hit rays.
// Entry = 16-byte header {env, agent*R + first ray, mask of agent-hit pixels, subsample | hit agent << 8} + per pixel two float4:
//   {texel rgb, 1-dot^2} and {hit point x, hit point y, static intensity, is-agent-hit}.
// One warp per entry (every agent-hit pixel group is an independent unit of work, so the whole machine is busy).
// The warp keeps the env's first 32 lights one per lane, each with the occluder last found for that light from
// around the agent that was hit (persistent across entries and steps in the workspace: a hint, re-verified by the
// exact test), reads only the occluder runs it visits straight from HBM/L2, and writes the final screen pixels and
// the pooled RGB observation of the group.
// ---------------------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128, 6) dyn_kernel(const __grid_constant__ KArgs k) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, tw = (gridDim.x * blockDim.x) >> 5;
    const int A = k.s.n_agents, AF = A * k.s.n_model, R = k.p.res;
    const int reserved = *reinterpret_cast<volatile int*>(k.dyn_ctrl);
    const int count = reserved < k.dyn_cap ? reserved : k.dyn_cap;
    const bool sorted = k.s.occ_lines != nullptr;
    unsigned dyn_rays = 0, dyn_iters = 0;
    for (int ei = wg; ei < count; ei += tw) {
        const unsigned char* e = k.dyn_entries + (size_t)ei * k.dyn_stride;
        const int4 hdr = *reinterpret_cast<const int4*>(e);
        const unsigned mask = (unsigned)hdr.z;
        if (!mask) continue;                                   // slot reserved by a chunk that fell back inline
        const int sub = hdr.w & 0xff, tgt = hdr.w >> 8;
        const int n = hdr.x, ar = hdr.y;                       // env; agent * R + first ray of the group
        const int av = ar / R, r = ar - av * R;
        const int64_t ag = (int64_t)n * A + av;
        const int64_t o0 = ag * R + r;
        const int L = __ldg(k.s.line_widths + n);
        const int nlights = __ldg(k.s.light_widths + n);
        const float* lt = k.s.lights + 3 * (int64_t)__ldg(k.s.light_starts + n);
        // lights one per lane, each with the occluder remembered for this (env, agent that was hit)
        int* cache = k.dyn_cache + ((size_t)n * A + (tgt < A ? tgt : 0)) * 32;
        LaneLight ll;
        ll.x = ll.y = ll.i = 0.f;
        ll.occ = cache[lane];
        if (lane < nlights) { ll.x = __ldg(lt + 3 * lane); ll.y = __ldg(lt + 3 * lane + 1); ll.i = __ldg(lt + 3 * lane + 2); }
        const int occ_before = ll.occ;
        const bool use_sorted = sorted && nlights <= 32;
        OccEnv oe;
        const float4* seg = reinterpret_cast<const float4*>(k.s.lines) + __ldg(k.s.line_starts + n);
        if (use_sorted) {
            const int W = L - AF;
            oe.occ = reinterpret_cast<const float4*>(k.s.occ_lines) + __ldg(k.s.occ_starts + n);
            oe.boxes = reinterpret_cast<const float4*>(k.s.occ_boxes) + __ldg(k.s.box_starts + n);
            oe.W = W; oe.run = k.s.occ_run; oe.nb = (W + oe.run - 1) / oe.run;
            oe.vmax = __ldg(k.s.occ_meta + 2 * n); oe.diam = __ldg(k.s.occ_meta + 2 * n + 1);
        } else if (ll.occ < AF || ll.occ >= L) ll.occ = -1;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < sub) {
            const float4* rec = reinterpret_cast<const float4*>(e + 16) + 2 * lane;
            ra = rec[0];
            rb = rec[1];
        }
        float intensity = rb.z;
        unsigned m = mask;
        while (m) {
            const int p = __ffs(m) - 1;
            m &= m - 1;
            const float cx = __shfl_sync(0xffffffffu, rb.x, p), cy = __shfl_sync(0xffffffffu, rb.y, p);
            const float v = use_sorted ? light_intensity_boxed<STATS>(oe, nlights, cx, cy, lane, ll, dyn_iters)
                                       : light_intensity_cached<STATS>(seg, L, AF, nlights, lt, cx, cy, lane, ll, dyn_iters);
            if (lane == p) intensity = v;
            if (STATS) dyn_rays++;
        }
        if (ll.occ != occ_before) cache[lane] = ll.occ;        // racy on purpose: any stored value is only a hint
        const float kk = fmul(ra.w, intensity);
        const float s0 = fmul(kk, ra.x), s1 = fmul(kk, ra.y), s2 = fmul(kk, ra.z);
        if (k.out.screen && lane < sub && ((mask >> lane) & 1u)) {
            float* sc = k.out.screen + 3 * (o0 + lane);
            sc[0] = s0; sc[1] = s1; sc[2] = s2;
        }
        if (k.has_obs && k.obs.rgb) {
            float v0 = lane < sub ? s0 : 0.f, v1 = lane < sub ? s1 : 0.f, v2 = lane < sub ? s2 : 0.f;
            for (int o = 1; o < sub; o <<= 1) {
                v0 = __fadd_rn(v0, __shfl_xor_sync(0xffffffffu, v0, o));
                v1 = __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                v2 = __fadd_rn(v2, __shfl_xor_sync(0xffffffffu, v2, o));
            }
            if (lane == 0) {
                const int Ro = R / sub, ro = r / sub;
                float* q = k.obs.rgb + ag * 3 * Ro + ro;
                q[0] = __fmul_rn(v0, k.inv_sub); q[Ro] = __fmul_rn(v1, k.inv_sub); q[2 * Ro] = __fmul_rn(v2, k.inv_sub);
            }
        }
    }
    if (STATS && k.stats && lane == 0) {
        atomicAdd(k.stats + STAT_DYN_RAYS, (unsigned long long)dyn_rays);
        atomicAdd(k.stats + STAT_DYN_ITERS, (unsigned long long)dyn_iters);
    }
    // the last CTA out re-arms the queue for the next step
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(k.dyn_ctrl + 1, 1) == (int)gridDim.x - 1) { k.dyn_ctrl[0] = 0; k.dyn_ctrl[1] = 0; }
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
    const Smem m = carve(smem_raw, k.seg_cap, nwarps, A);
    const int L = __ldg(k.s.line_widths + n);
    const int64_t g0 = __ldg(k.s.line_starts + n);
    if (tid == 0) {
        mbar_init(m.bar, 1);
        if (L > 0) {
            mbar_expect_tx(m.bar, (uint32_t)L * 16u);
            bulk_g2s(m.seg, k.s.lines + 4 * g0, (uint32_t)L * 16u, m.bar);
        }
    }
    __syncthreads();
    if (L > 0) mbar_wait(m.bar, 0);
    const int I = k.s.light_widths[n];
    const float* lt = k.s.lights + 3 * (int64_t)k.s.light_starts[n];
    for (int l = warp; l < L; l += nwarps) {
        const int w = __ldg(k.s.tex_widths + g0 + l);
        const int64_t ts = __ldg(k.s.tex_starts + g0 + l);
        const float4 s4 = m.seg[l];
        const float rw = rcp((float)w);
        for (int t = lane; t < w; t += 32) {
            const float loc = fmul(fadd((float)(unsigned)t, 0.5f), rw);                     // :278
            const float om = fsub(1.f, loc);
            const float Cx = ffma(s4.x, om, fmul(loc, s4.z)), Cy = ffma(s4.y, om, fmul(loc, s4.w));   // :279
            k.s.baked[ts + t] = light_intensity_thread(m.seg, AF, L, lt, I, Cx, Cy);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: C ABI
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static long long g_launches = 0;
static long long g_opt_nch = 0;          // 0 = auto
static long long g_opt_threads = 0;      // 0 = auto
static long long g_opt_skip_dyn = 0;     // debug
static long long g_opt_split = 0;        // 1: cast kernel + shade kernel (measured slower than one kernel: 186 vs 174 us)
static long long g_opt_two_phase = 0;    // 1: bin every (agent, segment) once into shared memory first (measured slower: the
                                         // records cost 70 KB per CTA, which halves residency)
static long long g_opt_variant = 0;      // experiment switches (see KArgs::variant)
static long long g_opt_split_step = 0;   // 1: msb_step launches physics and render separately
static long long g_opt_legacy = 0;       // 1: render with the line-order env_kernel instead of view_kernel
static long long g_opt_fused_step = 0;   // 1: msb_step runs physics and render in ONE kernel (slower: see DESIGN.md)
static unsigned long long* g_stats = nullptr;   // device counters, enabled by option "stats"

// Optional per-kernel timing with CUDA events on the launching stream (option "timing" = 1): bench.py uses it to
// report each kernel's live share of the step. Kinds: 0 physics, 1 render, 2 shade, 3 dyn, 4 fused step, 5 bake.
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
extern "C" int64_t msb_launch_count(void) { return g_launches; }

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
    if (!strcmp(name, "legacy_render")) { g_opt_legacy = value; return 0; }
    if (!strcmp(name, "split_step")) { g_opt_split_step = value; return 0; }
    if (!strcmp(name, "variant")) { g_opt_variant = value; return 0; }
    if (!strcmp(name, "timing")) {
        timing_flush();
        g_opt_timing = value;
        for (int i = 0; i < TK_KINDS; i++) { g_time_ms[i] = 0; g_time_n[i] = 0; }
        return 0;
    }
    if (!strcmp(name, "two_phase")) { g_opt_two_phase = value; return 0; }
    if (!strcmp(name, "split_render")) { g_opt_split = value; return 0; }
    if (!strcmp(name, "stats")) {
        if (value && !g_stats) {
            if (check(cudaMalloc(&g_stats, 8 * sizeof(unsigned long long)), "cudaMalloc(stats)")) return 1;
            return check(cudaMemset(g_stats, 0, 8 * sizeof(unsigned long long)), "cudaMemset(stats)");
        }
        if (!value && g_stats) { cudaFree(g_stats); g_stats = nullptr; }
        return 0;
    }
    if (!strcmp(name, "stats_reset")) {
        if (g_stats) return check(cudaMemset(g_stats, 0, 8 * sizeof(unsigned long long)), "cudaMemset(stats)");
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
        unsigned long long h[8];
        if (cudaMemcpy(h, g_stats, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (!strcmp(name, "stat_tests")) return (int64_t)h[STAT_TESTS];
        if (!strcmp(name, "stat_groups")) return (int64_t)h[STAT_GROUPS];
        if (!strcmp(name, "stat_dyn_rays")) return (int64_t)h[STAT_DYN_RAYS];
        if (!strcmp(name, "stat_dyn_iters")) return (int64_t)h[STAT_DYN_ITERS];
        if (!strcmp(name, "stat_replays")) return (int64_t)h[STAT_REPLAYS];
    }
    return -1;
}

static int validate(const msb_params* p, const msb_scenery* s) {
    if (!p || !s) return fail("%s", "null params/scenery");
    if (s->n_envs < 0 || s->n_agents < 1 || s->n_model < 0) return fail("%s", "bad scenery dimensions");
    if (s->max_lines < s->n_agents * s->n_model) return fail("%s", "scenery.max_lines is smaller than n_agents*n_model");
    if (s->max_lines > 14000) return fail("%s", "scene too large: more than 14000 segments in one environment");
    return 0;
}

template <int MODE>
static int launch_env(const KArgs& k, int nch, int threads, cudaStream_t st) {
    const size_t sm = smem_bytes(k.seg_cap, threads / 32, k.s.n_agents, (MODE & MODE_RENDER) && k.two_phase);
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
#define MSB_LAUNCH(N)                                                                                            \
    {                                                                                                            \
        auto fn = k.split_render ? (k.stats ? env_kernel<MODE, N, true, true> : env_kernel<MODE, N, false, true>)    \
                                 : (k.stats ? env_kernel<MODE, N, true, false> : env_kernel<MODE, N, false, false>); \
        if (sm > 48 * 1024 && check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), \
                                    "cudaFuncSetAttribute"))                                                     \
            return 1;                                                                                            \
        fn<<<k.s.n_envs, threads, sm, st>>>(k);                                                                  \
    }
    {
        TimedLaunch timed(MODE == MODE_PHYSICS ? TK_PHYSICS : (MODE == MODE_RENDER ? TK_RENDER : TK_STEP), st);
        switch (nch) {
            case 1: MSB_LAUNCH(1); break;
            case 2: MSB_LAUNCH(2); break;
            default: MSB_LAUNCH(4); break;
        }
    }
#undef MSB_LAUNCH
    g_launches++;
    return check(cudaGetLastError(), "kernel launch");
}

static bool use_view(const KArgs& k) {
    return !g_opt_legacy && k.s.occ_lines && k.s.occ_rec && !k.split_render && !k.two_phase;
}

static int launch_view(const KArgs& k, bool phys, int nch, int threads, cudaStream_t st) {
    const size_t sm = vsmem_bytes(k.wcap, threads / 32, k.s.n_agents, k.s.n_agents * k.s.n_model);
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
#define MSB_LAUNCH(N)                                                                                            \
    {                                                                                                            \
        auto fn = phys ? (k.stats ? view_kernel<N, true, true> : view_kernel<N, true, false>)                    \
                       : (k.stats ? view_kernel<N, false, true> : view_kernel<N, false, false>);                 \
        if (sm > 48 * 1024 && check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), \
                                    "cudaFuncSetAttribute"))                                                     \
            return 1;                                                                                            \
        fn<<<k.s.n_envs, threads, sm, st>>>(k);                                                                  \
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
    return check(cudaGetLastError(), "kernel launch");
}

static void plan_render(const msb_params* p, const msb_scenery* s, KArgs& k, int* nch, int* rb, int* threads) {
    const int chunks = (p->res + 31) / 32;
    // split render needs the four scalar Render outputs as the hand-over buffers
    k.split_render = g_opt_split && k.out.indices && k.out.locations && k.out.dots && k.out.distances;
    // two-phase render when the per-(agent, segment) records leave room for at least two CTAs per SM
    const size_t rec = (size_t)2 * s->n_agents * k.seg_cap * 16;
    k.two_phase = (g_opt_two_phase != 0) && rec + (size_t)k.seg_cap * 18 + 4096 <= 100 * 1024;
    if (k.two_phase) {
        int n1 = (g_opt_nch == 1 || g_opt_nch == 2 || g_opt_nch == 4) ? (int)g_opt_nch : 1;
        while (n1 > chunks) n1 >>= 1;
        *nch = n1;
        *rb = (chunks + n1 - 1) / n1;
        int t = 32 * s->n_agents * (*rb);
        if (t < 64) t = 64;
        if (t > 256) t = 256;
        if (g_opt_threads >= 32 && g_opt_threads <= 256) t = (int)(g_opt_threads / 32) * 32;
        *threads = t;
        return;
    }
    int n = 4;
    if (g_opt_nch == 1 || g_opt_nch == 2 || g_opt_nch == 4) n = (int)g_opt_nch;
    else {
        // largest chunk count per warp that still leaves the env's CTA at least 4 warps of work
        while (n > 1 && (s->n_agents * ((chunks + n - 1) / n) < 4 || n > chunks)) n >>= 1;
    }
    *nch = n;
    *rb = (chunks + n - 1) / n;
    int t = 32 * s->n_agents * (*rb);
    if (t < 64) t = 64;
    if (t > 256) t = 256;
    if (g_opt_threads >= 32 && g_opt_threads <= 256) t = (int)(g_opt_threads / 32) * 32;
    *threads = t;
}

static void fill(KArgs& k, const msb_params* p, const msb_scenery* s, const msb_agents* a) {
    memset(&k, 0, sizeof(k));
    k.p = *p;
    k.s = *s;
    if (a) k.a = *a;
    k.seg_cap = s->max_lines > 0 ? s->max_lines : 1;
    if (k.s.occ_run != VRUN) { k.s.occ_run = VRUN; k.s.occ_lines = nullptr; }      // the table's runs must be 16 long
    if (!k.s.occ_lines || !k.s.occ_boxes || !k.s.box_starts || !k.s.occ_starts) { k.s.occ_lines = nullptr; k.s.occ_rec = nullptr; }
    {
        const int w = s->max_lines - s->n_agents * s->n_model;
        k.wcap = w > 0 ? ((w + VRUN - 1) / VRUN) * VRUN : VRUN;
    }
    k.inv_fps = 1.0f / p->fps;
    k.ray_blocks = 1;
    k.stats = g_stats;
    k.debug_skip_dyn = (int32_t)g_opt_skip_dyn;
    k.variant = (int32_t)g_opt_variant;
    k.bin_kappa = (float)p->res / (2.f * p->half_screen);
    k.bin_rmid = 0.5f * ((float)p->res - 1.f);
    k.bin_xclip = 0.5f * p->agent_radius / sqrtf(1.f + p->half_screen * p->half_screen);
}

extern "C" int msb_physics(const msb_params* p, const msb_scenery* s, const msb_agents* a, float* progress,
                           void* cuda_stream) {
    if (validate(p, s)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_physics: null agents");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    k.progress = progress;
    int threads = 128;
    if (g_opt_threads >= 32 && g_opt_threads <= 256) threads = (int)(g_opt_threads / 32) * 32;
    return launch_env<MODE_PHYSICS>(k, 1, threads, (cudaStream_t)cuda_stream);
}

// workspace layout: int ctrl[4] (16 bytes) | occluder cache int[N][A][32] | entries of (16 + 32*subsample) bytes
static int64_t cache_bytes(const msb_scenery* s) { return (int64_t)s->n_envs * s->n_agents * 32 * 4; }

static int set_workspace(KArgs& k, const msb_workspace* ws) {
    k.dyn_ctrl = nullptr;
    k.dyn_entries = nullptr;
    k.dyn_cache = nullptr;
    k.dyn_cap = 0;
    const int sub = k.has_obs ? k.obs.subsample : 1;
    k.dyn_stride = 16 + 32 * sub;
    if (!ws || !ws->ptr) return 0;
    if (((uintptr_t)ws->ptr & 15) != 0) return fail("%s", "workspace must be 16-byte aligned");
    const int64_t head = 16 + cache_bytes(&k.s);
    const int64_t cap = (ws->bytes - head) / k.dyn_stride;
    if (cap < 1) return 0;
    k.dyn_ctrl = reinterpret_cast<int*>(ws->ptr);
    k.dyn_cache = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws->ptr) + 16);
    k.dyn_entries = reinterpret_cast<unsigned char*>(ws->ptr) + head;
    k.dyn_cap = cap > 0x7fffffff ? 0x7fffffff : (int32_t)cap;
    return 0;
}

static int launch_shade(const KArgs& k, cudaStream_t st) {
    if (!k.split_render) return 0;
    const int64_t warps = (int64_t)k.s.n_envs * k.s.n_agents * ((k.p.res + 31) / 32);
    const int64_t blocks = (warps + 7) / 8;
    {
        TimedLaunch timed(TK_SHADE, st);
        if (k.stats) shade_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(k);
        else shade_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(k);
    }
    g_launches++;
    return check(cudaGetLastError(), "shade_kernel launch");
}

static int launch_dyn(const KArgs& k, cudaStream_t st) {
    if (!k.dyn_entries) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 12;                      // 48 warps per SM, each striding over the queue
    {
        TimedLaunch timed(TK_DYN, st);
        if (k.stats) dyn_kernel<true><<<grid, 128, 0, st>>>(k);
        else dyn_kernel<false><<<grid, 128, 0, st>>>(k);
    }
    g_launches++;
    return check(cudaGetLastError(), "dyn_kernel launch");
}

extern "C" int64_t msb_workspace_bytes(const msb_params* p, const msb_scenery* s, int32_t subsample) {
    if (!p || !s || subsample < 1) return 0;
    const int64_t rays = (int64_t)s->n_envs * s->n_agents * p->res;
    const int64_t groups = rays / subsample;
    // room for a quarter of all pixel groups to contain an agent-hit ray (overflow falls back to inline, still exact)
    int64_t cap = groups / 4;
    if (cap < 65536) cap = groups < 65536 ? groups : 65536;
    return 16 + cache_bytes(s) + cap * (16 + 32 * (int64_t)subsample);
}

static void set_obs(KArgs& k, const msb_obs_out* obs) {
    if (!obs) return;
    k.obs = *obs;
    k.has_obs = 1;
    k.inv_max_depth = 1.0f / obs->max_depth;
    k.inv_speed = 1.0f / obs->speed_scale;
    k.inv_ang = 1.0f / obs->ang_scale;
    k.inv_sub = 1.0f / (float)obs->subsample;
}

static int check_obs(const msb_params* p, const msb_obs_out* obs) {
    if (!obs) return 0;
    const int sub = obs->subsample;
    if (sub < 1 || sub > 32 || (sub & (sub - 1)) || p->res % sub) return fail("%s", "obs.subsample must be a power of two <= 32 dividing res");
    return 0;
}

extern "C" int msb_render(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_render_out* out,
                          const msb_obs_out* obs, const msb_workspace* ws, void* cuda_stream) {
    if (validate(p, s) || check_obs(p, obs)) return 1;
    if (!a || !a->angles || !a->positions) return fail("%s", "msb_render: null agents");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    if (out) k.out = *out;
    set_obs(k, obs);
    if (set_workspace(k, ws)) return 1;
    int nch, rb, threads;
    plan_render(p, s, k, &nch, &rb, &threads);
    k.ray_blocks = rb;
    if (use_view(k)) {
        if (launch_view(k, false, nch, threads, (cudaStream_t)cuda_stream)) return 1;
    } else if (launch_env<MODE_RENDER>(k, nch, threads, (cudaStream_t)cuda_stream)) return 1;
    if (launch_shade(k, (cudaStream_t)cuda_stream)) return 1;
    return launch_dyn(k, (cudaStream_t)cuda_stream);
}

extern "C" int msb_step(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv,
                        float* progress, const msb_render_out* out, const msb_obs_out* obs, const msb_workspace* ws,
                        void* cuda_stream) {
    if (validate(p, s) || check_obs(p, obs)) return 1;
    if (!a || !a->angles || !a->positions || !a->angvelocity || !a->velocity) return fail("%s", "msb_step: null agents");
    if (mv && !mv->actions) return fail("%s", "msb_step: movement without actions");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, a);
    k.progress = progress;
    if (out) k.out = *out;
    set_obs(k, obs);
    if (mv) {
        k.mv = *mv;
        k.has_mv = 1;
        k.mv_keep = (float)(1.0 - (double)mv->decay);
        k.mv_dv = (float)((double)mv->accel / (double)p->fps);
        k.mv_dw = (float)((double)mv->ang_accel / (double)p->fps);
    }
    if (set_workspace(k, ws)) return 1;
    int nch, rb, threads;
    plan_render(p, s, k, &nch, &rb, &threads);
    k.ray_blocks = rb;
    if (use_view(k) && !g_opt_split_step) {
        // the whole tick in ONE kernel: movement + physics over the staged table, draw, render, heads
        if (launch_view(k, true, nch, threads, (cudaStream_t)cuda_stream)) return 1;
    } else if (g_opt_fused_step) {
        k.split_render = 0;
        if (launch_env<MODE_STEP>(k, nch, threads, (cudaStream_t)cuda_stream)) return 1;
    } else {
        int pthreads = 128;
        if (g_opt_threads >= 32 && g_opt_threads <= 256) pthreads = (int)(g_opt_threads / 32) * 32;
        if (launch_env<MODE_PHYSICS>(k, 1, pthreads, (cudaStream_t)cuda_stream)) return 1;
        if (use_view(k)) {
            if (launch_view(k, false, nch, threads, (cudaStream_t)cuda_stream)) return 1;
        } else if (launch_env<MODE_RENDER>(k, nch, threads, (cudaStream_t)cuda_stream)) return 1;
    }
    if (launch_shade(k, (cudaStream_t)cuda_stream)) return 1;
    return launch_dyn(k, (cudaStream_t)cuda_stream);
}

extern "C" int msb_bake(const msb_params* p, const msb_scenery* s, void* cuda_stream) {
    if (validate(p, s)) return 1;
    if (!s->baked) return fail("%s", "msb_bake: null baked");
    if (s->n_envs == 0) return 0;
    KArgs k;
    fill(k, p, s, nullptr);
    const int threads = 256;
    const size_t sm = smem_bytes(k.seg_cap, threads / 32, s->n_agents);
    if (sm > 227 * 1024) return fail("%s", "scene too large: an env's segments do not fit in shared memory (227 KB)");
    if (sm > 48 * 1024 &&
        check(cudaFuncSetAttribute(bake_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "cudaFuncSetAttribute"))
        return 1;
    bake_kernel<<<s->n_envs, threads, sm, (cudaStream_t)cuda_stream>>>(k);
    g_launches++;
    return check(cudaGetLastError(), "bake launch");
}
