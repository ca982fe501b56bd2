/* megastep_b200.h — C ABI of the B200-native simulation core (libmegastep_b200.so).
 *
 * This is the drop-in boundary for the reference's per-step hot path. The reference exposes that path as a
 * pybind11/ATen extension (`megastepcuda`, megastep/src/wrappers.cpp:30-172); the entry points below are what a
 * binding for it binds, with every at::Tensor flattened to a raw DEVICE pointer plus sizes, and the reference's
 * process-global `initialize()` constants (megastep/src/kernels.cu:12-27) turned into an explicit per-call params
 * struct. No torch types, no allocation: the caller owns every buffer (scratch included); kernels are enqueued on the
 * caller's stream and return without synchronising (as the reference: kernels.cu:30-32, no device sync anywhere).
 * Every entry point takes all of its state through its arguments and may be called from several threads (on different
 * streams, each with its own workspace) — with three process-global exceptions, none of which changes results:
 * msb_set_option / msb_get_option (tuning and diagnostics switches, the per-kernel timing ring and the stats counters:
 * set them from one thread, while nothing else is launching), msb_launch_count (an atomic counter) and msb_last_error
 * (thread-local).
 *
 * All functions return 0 on success, non-zero on failure (msb_last_error() describes it). Pointer arguments are
 * device pointers unless marked [host]. All float data is fp32; all index data int32 unless stated (texel
 * offsets are int64 so scenes beyond 2^31 texels work, which the reference's 32-bit accessors —
 * megastep/src/common.h:30,43 — cannot address).
 *
 * INTEGRATION.md shows the binding a reference maintainer would add on top of this header.
 */
#ifndef MEGASTEP_B200_H
#define MEGASTEP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB_ABI_VERSION 5

/* Replaces initialize(agent_radius, res, fov, fps) — megastep/src/wrappers.cpp:53, kernels.cu:18-27. */
typedef struct msb_params {
    int32_t res;            /* R: rays (pixels) per agent; any R >= 1 (the reference caps at 1024, core.py:61-64) */
    float agent_radius;     /* collision radius and near plane; megastep/core.py:14 */
    float half_screen;      /* tanf(pi/180*fov/2.) — filled by msb_params_init exactly as kernels.cu:22 does */
    float fps;              /* steps per second */
    float fov;              /* degrees, kept for reference */
    int32_t reserved[3];
} msb_params;

/* The Scenery struct of megastep/src/common.h:185-214, as raw ragged arrays.
 *   lines   : ragged per env  — vals (sum L, 2, 2) = one float4 {ax, ay, bx, by} per segment; the first
 *             n_agents*n_model lines of every env are the agents' model lines (render() rewrites them).
 *   lights  : ragged per env  — vals (sum I, 3) = {x, y, intensity}
 *   textures: ragged per LINE — vals (sum T, 3) linear RGB; widths (sum L) texels per line
 *   baked   : (sum T) baked light per texel, same raggedness as textures
 *   model   : (n_model, 2, 2) the agent outline in agent-local coordinates
 */
typedef struct msb_scenery {
    int32_t n_envs;             /* N */
    int32_t n_agents;           /* A: agents per env */
    int32_t n_model;            /* F: lines in the agent model */
    int32_t max_lines;          /* max over envs of line_widths (shared-memory sizing) */
    int32_t max_lights;         /* max over envs of light_widths */
    int32_t occ_run;            /* segments per run of the spatial table: 16 (only read when occ_lines is set) */
    float* lines;               /* (sum L, 4) — render()/step() write the agents' lines in place */
    const int32_t* line_widths; /* (N) */
    const int32_t* line_starts; /* (N) exclusive prefix sum of line_widths */
    const float* lights;        /* (sum I, 3) */
    const int32_t* light_widths;/* (N) */
    const int32_t* light_starts;/* (N) */
    const float* textures;      /* (sum T, 3) */
    const int32_t* tex_widths;  /* (sum L) */
    const int64_t* tex_starts;  /* (sum L) exclusive prefix sum of tex_widths, 64-bit */
    float* baked;               /* (sum T) — written by msb_bake, read by msb_render */
    const float* model;         /* (F, 4) */
    int64_t n_lines;            /* sum L */
    int64_t n_texels;           /* sum T */
    /* The spatial table, required by msb_physics / msb_render / msb_step: a copy of every env's STATIC segments packed
     * sort-tile-recursive into runs of occ_run = 16 with one bounding box per run; each env's rows are padded to a
     * whole number of runs (so every env's block is 16-byte aligned for bulk copies). The caller allocates the arrays
     * and fills box_starts (exclusive prefix sum of nb); msb_build_table fills the rest, once per scenery. Used to skip
     * work only: shadow tests and collisions are order-free, and render() restores the reference's line-order rule. */
    const float* occ_lines;     /* (16 * sum nb, 4) sorted static segments; nb = ceil((line_widths[n] - A*F) / 16) per env */
    const int32_t* occ_starts;  /* (N) start of env n's rows in occ_lines (= 16 * box_starts[n]) */
    const float* occ_boxes;     /* (sum nb, 4) {xmin, ymin, xmax, ymax} of each run */
    const int32_t* box_starts;  /* (N) start of env n's rows in occ_boxes */
    const float* occ_meta;      /* (N, 2) per env: longest static segment extent (max |dx|,|dy|), extent of the env */
    const int32_t* occ_rec;     /* (16 * sum nb, 4) per sorted row: {tex_starts lo, hi, tex_widths, line index within its env};
                                 * padding rows have line index -1 */
    /* Optional light-visibility grid (vis NULL to disable; needs the spatial table): 0.25 m cells over each env's static
     * geometry. Bit i of a cell's word is set when light i (< 32) of the env is CERTAINLY unoccluded from every point
     * of the cell — no static segment comes near any segment light -> point, by a margin that covers the rounding of
     * the reference's intersect() — so the dynamic light of a ray that hit an agent needs no shadow test for it.
     * Unset bits promise nothing. Allocated by the caller, filled once per scenery by msb_build_visibility. */
    uint32_t* vis;              /* (sum gx*gy) */
    const int64_t* vis_starts;  /* (N) first cell of env n */
    const float* vis_meta;      /* (N, 4) {x0, y0, gx, gy}: the grid's origin in metres and its dimensions (whole numbers) */
    /* Optional launch order (NULL: env n is CTA n): a permutation of 0..N-1, costliest envs first, so that the grid's
     * last CTAs are its cheapest. Speed only. */
    const int32_t* env_order;   /* (N) */
} msb_scenery;

/* The Agents struct of megastep/src/common.h:162-177. Updated in place by msb_physics. */
typedef struct msb_agents {
    float* angles;       /* (N, A)    degrees in [-180, 180) */
    float* positions;    /* (N, A, 2) metres */
    float* angvelocity;  /* (N, A)    degrees / s */
    float* velocity;     /* (N, A, 2) metres / s */
} msb_agents;

/* The Render struct of megastep/src/common.h:216-222. Any pointer may be NULL to skip that output. */
typedef struct msb_render_out {
    int32_t* indices;    /* (N, A, R)   line index of the nearest hit, -1 for none */
    float* locations;    /* (N, A, R)   position along the hit line in [0, 1], NaN for none */
    float* dots;         /* (N, A, R)   cosine between ray and line, NaN for none */
    float* distances;    /* (N, A, R)   metres to the hit, +inf for none */
    float* screen;       /* (N, A, R, 3) linear RGB */
} msb_render_out;

/* Optional fused observation heads (reference: megastep/modules.py:170-184 Depth, :211-224 RGB, :263-270 IMU),
 * produced by msb_render / msb_step in the same pass when the pointers are non-NULL. */
typedef struct msb_obs_out {
    float* rgb;          /* (N, A, 3, R/subsample)  mean over `subsample` adjacent pixels */
    float* depth;        /* (N, A, R/subsample)     1 - clamp((distance - agent_radius)/max_depth, 0, 1), mean-pooled */
    float* imu;          /* (N, A, 3)               {angvelocity/ang_scale, local-frame velocity/speed_scale} */
    int32_t subsample;   /* >= 1, divides R */
    float max_depth;     /* modules.py:147 default 10 */
    float speed_scale;   /* modules.py:242 default 10 */
    float ang_scale;     /* modules.py:242 default 360 */
} msb_obs_out;

/* MomentumMovement (megastep/modules.py:68-118), fused into msb_step ahead of the physics. */
typedef struct msb_movement {
    const int32_t* actions; /* (N, A) in [0, 7): 0 noop, 1 +y, 2 -y, 3 +x, 4 -x (agent-local metres), 5 turn +, 6 turn - (modules.py:95-96) */
    float accel;            /* m/s^2, default 5 */
    float ang_accel;        /* deg/s^2, default 180 */
    float decay;            /* default 0.125 */
} msb_movement;

/* Scratch for the second pass that lights the rays which hit agents (their light is dynamic: kernels.cu:434-436).
 * Caller-owned device memory, 16-byte aligned, zero-filled once before first use; sized by msb_workspace_bytes
 * (any size works — what does not fit is resolved inline by the first pass, same results, longer tail). Holds the
 * queue of agent-hit pixel windows (filled by the first pass while the second — a programmatic dependent launch —
 * already drains it), its counters and the per-(env, agent) occluder hints, which persist from step to step.
 * One workspace per concurrent stream. NULL disables the second pass. */
typedef struct msb_workspace {
    void* ptr;
    int64_t bytes;
} msb_workspace;

int msb_abi_version(void);
const char* msb_last_error(void);

/* [host] Fills *p. half_screen = tanf(CUDART_PI_F/180.f*fov/2.) evaluated as the reference does. fov must be < 180. */
int msb_params_init(msb_params* p, float agent_radius, int32_t res, float fov, float fps);

/* bake(scenery) — megastep/src/wrappers.cpp:61, kernels.cu:270-293. Writes scenery->baked for every texel. */
int msb_bake(const msb_params* p, const msb_scenery* s, void* cuda_stream);

/* Fills the spatial table (s->occ_lines, occ_rec, occ_boxes, occ_meta, occ_starts — written through the const pointers)
 * from s->lines, the texture metadata and s->box_starts. One-off; rerun it if the static lines change. */
int msb_build_table(const msb_scenery* s, void* cuda_stream);

/* Fills s->vis (see msb_scenery) from the static segments (spatial table) and lights. One-off, like msb_bake. */
int msb_build_visibility(const msb_scenery* s, void* cuda_stream);

/* physics(scenery, agents) -> Physics{progress} — wrappers.cpp:69, kernels.cu:179-230.
 * progress: (N, A) out. Agents are advanced in place (positions, angles) and stopped where progress < 1. */
int msb_physics(const msb_params* p, const msb_scenery* s, const msb_agents* a, float* progress, void* cuda_stream);

/* MomentumMovement.__call__ — megastep/modules.py:106-118 — as ONE launch: the velocities decay and take the chosen
 * actions' impulses, then physics() as above (the reference: ~10 PyTorch launches, then physics). */
int msb_move(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv, float* progress,
             void* cuda_stream);

/* render(scenery, agents) -> Render — wrappers.cpp:82, kernels.cu:297-475. Also rewrites the agents' model
 * lines inside s->lines (the reference's draw_kernel side effect). obs may be NULL. */
int msb_render(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_render_out* out,
               const msb_obs_out* obs, const msb_workspace* ws, void* cuda_stream);

/* One whole environment tick: MomentumMovement + physics (one kernel), render + observation heads (one kernel) and —
 * when a workspace is given — the small second pass that lights agent-hit rays. No host round trip in between.
 * mv may be NULL (velocities are then taken as already set, i.e. plain physics+render). out/obs as msb_render. */
int msb_step(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv,
             float* progress, const msb_render_out* out, const msb_obs_out* obs, const msb_workspace* ws,
             void* cuda_stream);

/* The same tick captured once as a CUDA graph (its programmatic dependencies included) and driven from the host in
 * ONE call: msb_step_graph_run uploads the actions from pinned host memory (NULL: already on the device), replays
 * the graph, downloads `progress` into pinned host memory (NULL: leave it on the device) and, if `sync`, waits for the
 * stream. All pointers given to msb_step_graph_create are baked in and must outlive the graph. The step's kernels
 * must have run at least once before (msb_step), and per-kernel timing must be off. */
typedef struct msb_graph msb_graph;
int msb_step_graph_create(const msb_params* p, const msb_scenery* s, const msb_agents* a, const msb_movement* mv,
                          float* progress, const msb_render_out* out, const msb_obs_out* obs, const msb_workspace* ws,
                          msb_graph** handle);
int msb_step_graph_run(msb_graph* g, const int32_t* actions_host, float* progress_host, void* cuda_stream, int32_t sync);
int msb_step_graph_destroy(msb_graph* g);

/* [host] Recommended workspace size in bytes for this scene and observation subsample (1 when obs is NULL). */
int64_t msb_workspace_bytes(const msb_params* p, const msb_scenery* s, int32_t subsample);

/* Tuning/diagnostics: selects kernel variants (0 = default). Affects speed only, never results. Process-global and NOT
 * thread-safe (see the top of this header): "timing" records CUDA events around every kernel on the launching stream,
 * "stats" allocates device counters the kernels add to. */
int msb_set_option(const char* name, int64_t value);
int64_t msb_get_option(const char* name);

/* ---- Environment rules on the device (the game logic of the reference's demo envs). Each is one small
 * launch on the caller's stream, no host round trip; `indices` / `locations` are msb_render_out tensors. ----------------- */

/* Explorer's reward bookkeeping — megastep/demo/envs/explorer.py:34-58. `seen`: one bit per texel of the whole scene
 * (ceil(n_texels / 32) words, zeroed once); every ray's texel (min(floor(width * location), width - 1) on the line it hit)
 * is marked, and the number of bits NEWLY set per env is added to potential[n] (the env's seen-texel count, what the
 * reference recomputes with a scatter_add over every texel each step) and to gained[n] (zero it before the call). */
int msb_env_ledger_mark(const msb_scenery* s, const int32_t* indices, const float* locations, int32_t n_agents, int32_t res,
                        uint32_t* seen, int32_t* potential, int32_t* gained, void* cuda_stream);

/* explorer.py:73-77: envs with reset[n] != 0 forget what they have seen (their bit range is cleared, potential[n] = 0). */
int msb_env_ledger_clear(const msb_scenery* s, const uint8_t* reset, uint32_t* seen, int32_t* potential, void* cuda_stream);

/* Deathmatch's crosshair rule and its consequences — megastep/demo/envs/deathmatch.py:54-72, 75-80. For every agent: the
 * agents whose model shows at the centre ray of one of its two middle pooled pixels (pooling = `subsample`) are hit.
 * matchings (N, A, A) uint8 [shooter][target]; hits (N, A) = targets hit by each agent; damage += .05 hits;
 * health += -.05 (times hit + outside the floorplan by more than `clearance`) - .001. bounds (N, 2) as the reference's. */
int msb_env_shoot(const msb_scenery* s, const msb_agents* a, const int32_t* indices, int32_t res, int32_t subsample,
                  const float* bounds, float clearance, uint8_t* matchings, float* hits, float* health, float* damage,
                  void* cuda_stream);

/* RandomSpawns — megastep/modules.py:312-326 — without its nonzero() (a device-to-host sync per step): agents with
 * reset[n][a] != 0 move to spawn `choices[n][a] % n_spawns` (or, choices NULL, to one drawn by a counter-based hash of
 * (seed, tick, agent)) of their precomputed spawn_positions (N, A, n_spawns, 2) / spawn_angles (N, A, n_spawns), with
 * their velocities zeroed. */
int msb_env_respawn(const msb_scenery* s, const msb_agents* a, const uint8_t* reset, const float* spawn_positions,
                    const float* spawn_angles, int32_t n_spawns, uint32_t seed, uint32_t tick, const int32_t* choices,
                    void* cuda_stream);

/* The three observation heads of a batch (msb_obs_out's rgb (N, A, 3, ro), depth (N, A, ro), imu (N, A, 3)) packed into
 * one row per env, [rgb | depth | imu], in one pass: the send buffer of the multi-GPU observation all-gather
 * (megastep_b200/sharding.py). mode 0: fp32 rows; 1: fp16 rows; 2: rgb and depth as uint8 = clamp(round(255 x), 0, 255),
 * the imu as fp16 starting at byte `imu_offset` of the row. `row_bytes` is the row pitch. */
int msb_pack_obs(const float* rgb, const float* depth, const float* imu, int64_t n_envs, int32_t n_agents, int32_t ro,
                 void* rows, int64_t row_bytes, int32_t mode, int32_t imu_offset, void* cuda_stream);

/* Number of kernels launched by this library since load (bench.py's `gpu_launches`). */
int64_t msb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MEGASTEP_B200_H */
