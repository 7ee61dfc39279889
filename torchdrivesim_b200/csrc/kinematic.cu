// Kinematic state transition, forward and backward (one thread per agent, float4 state I/O).
//
// Reference semantics: KinematicBicycle.step (torchdrivesim/kinematic.py:462-477),
// BicycleNoReversing.step (:509-523); the per-agent `model` id replaces the boolean-mask
// split/merge of CompoundKinematicModel.step (:197-201), which costs a host sync per step.
// The unicycle is defined by this build (the reference only names it, README.md:16).
// HBM traffic per agent: 16 B state in + 8 B action + 4 B lr (+4 B model) + 16 B state out.
#include "tds_common.cuh"

namespace {

struct KinIn {
    float x, y, psi, v, a0, a1, a2, a3, lr;
    int model;
};

__device__ __forceinline__ bool is_free_model(int model) { return model == TDS_MODEL_SIMPLE || model == TDS_MODEL_ORIENTED; }

__device__ __forceinline__ KinIn load_agent(const float* state, const float* action, int action_dim, const float* lr,
                                            const int32_t* model, int uniform_model, int64_t i) {
    const float4 s = reinterpret_cast<const float4*>(state)[i];
    KinIn k;
    k.x = s.x; k.y = s.y; k.psi = s.z; k.v = s.w;
    k.a2 = k.a3 = 0.0f;
    if (action_dim == 4) {
        const float4 a = reinterpret_cast<const float4*>(action)[i];
        k.a0 = a.x; k.a1 = a.y; k.a2 = a.z; k.a3 = a.w;
    } else {
        const float2 a = reinterpret_cast<const float2*>(action)[i];
        k.a0 = a.x; k.a1 = a.y;
    }
    k.model = model ? model[i] : uniform_model;
    k.lr = (k.model >= TDS_MODEL_UNICYCLE || lr == nullptr) ? 1.0f : lr[i];
    return k;
}

// de-normalised controls shared by forward and backward
struct Controls {
    float acc, ang;      // acceleration, steering (bicycle) or yaw rate (unicycle) after all flips
    bool reversing;      // BicycleNoReversing clamp active
};

__device__ __forceinline__ Controls controls(const KinIn& k, const tds_kinematic_params_t& p) {
    Controls c;
    const float ang_scale = k.model == TDS_MODEL_UNICYCLE ? p.max_yaw_rate : p.max_steering;
    c.acc = k.a0 * p.max_acceleration;
    c.ang = k.a1 * ang_scale;
    c.reversing = false;
    if (k.model == TDS_MODEL_BICYCLE_NO_REVERSING) {
        c.reversing = (k.v + c.acc * p.dt) < 0.0f;
        if (c.reversing) c.acc = (-k.v) / p.dt;
        // the reference re-normalises and de-normalises the modified action (kinematic.py:521-523)
        c.acc = (c.acc / p.max_acceleration) * p.max_acceleration;
        c.ang = (c.ang / ang_scale) * ang_scale;
    }
    if (p.left_handed) c.ang = -c.ang;
    return c;
}

__global__ void __launch_bounds__(256) kin_fwd_kernel(const float* __restrict__ state, const float* __restrict__ action,
                                                      int action_dim, const float* __restrict__ lr,
                                                      const int32_t* __restrict__ model, int uniform_model, int64_t n,
                                                      tds_kinematic_params_t p, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KinIn k = load_agent(state, action, action_dim, lr, model, uniform_model, i);
    if (is_free_model(k.model)) {
        // state += (action * [max_dx, max_dx, max_dpsi, max_dv]) * dt; ORIENTED rotates the xy action by psi first
        float ax = k.a0, ay = k.a1;
        if (k.model == TDS_MODEL_ORIENTED) {
            float s, co;
            tds::sincos_cr(k.psi, s, co);
            ax = co * k.a0 + (-s) * k.a1;
            ay = s * k.a0 + co * k.a1;
        }
        float4 o;
        o.x = k.x + (ax * p.max_dx) * p.dt;
        o.y = k.y + (ay * p.max_dx) * p.dt;
        o.z = k.psi + (k.a2 * p.max_dpsi) * p.dt;
        o.w = k.v + (k.a3 * p.max_dv) * p.dt;
        reinterpret_cast<float4*>(out)[i] = o;
        return;
    }
    const Controls c = controls(k, p);
    const float v = k.v + c.acc * p.dt;
    float4 o;
    if (k.model == TDS_MODEL_UNICYCLE) {
        float s, co;
        tds::sincos_cr(k.psi, s, co);
        o.x = k.x + (v * co) * p.dt;
        o.y = k.y + (v * s) * p.dt;
        o.z = k.psi + c.ang * p.dt;
    } else {
        float s, co, sb, cb;
        tds::sincos_cr(k.psi + c.ang, s, co);
        tds::sincos_cr(c.ang, sb, cb);
        o.x = k.x + (v * co) * p.dt;
        o.y = k.y + (v * s) * p.dt;
        o.z = k.psi + ((v / k.lr) * sb) * p.dt;
    }
    o.w = v;
    reinterpret_cast<float4*>(out)[i] = o;
}

__device__ __forceinline__ void store_action_grad(float* g_action, int action_dim, int64_t i, float g0, float g1, float g2,
                                                  float g3) {
    if (!g_action) return;
    if (action_dim == 4) reinterpret_cast<float4*>(g_action)[i] = make_float4(g0, g1, g2, g3);
    else reinterpret_cast<float2*>(g_action)[i] = make_float2(g0, g1);
}

__global__ void __launch_bounds__(256) kin_bwd_kernel(const float* __restrict__ state, const float* __restrict__ action,
                                                      int action_dim, const float* __restrict__ lr,
                                                      const int32_t* __restrict__ model, int uniform_model, int64_t n,
                                                      tds_kinematic_params_t p, const float* __restrict__ grad_out,
                                                      float* __restrict__ g_state, float* __restrict__ g_action,
                                                      float* __restrict__ g_lr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KinIn k = load_agent(state, action, action_dim, lr, model, uniform_model, i);
    const float4 g = reinterpret_cast<const float4*>(grad_out)[i];   // d/d(x', y', psi', v')
    if (is_free_model(k.model)) {
        const float sx = p.max_dx * p.dt;
        float gpsi = g.z, ga0 = g.x * sx, ga1 = g.y * sx;
        if (k.model == TDS_MODEL_ORIENTED) {
            float s, co;
            tds::sincos_cr(k.psi, s, co);
            const float xr = co * k.a0 - s * k.a1, yr = s * k.a0 + co * k.a1;
            gpsi += (g.y * xr - g.x * yr) * sx;
            ga0 = (g.x * co + g.y * s) * sx;
            ga1 = (g.y * co - g.x * s) * sx;
        }
        if (g_state) reinterpret_cast<float4*>(g_state)[i] = make_float4(g.x, g.y, gpsi, g.w);
        store_action_grad(g_action, action_dim, i, ga0, ga1, g.z * p.max_dpsi * p.dt, g.w * p.max_dv * p.dt);
        if (g_lr) g_lr[i] = 0.0f;
        return;
    }
    const Controls c = controls(k, p);
    const float v = k.v + c.acc * p.dt;
    const float dt = p.dt;
    float gv, gpsi, gang, glr = 0.0f;
    if (k.model == TDS_MODEL_UNICYCLE) {
        float s, co;
        tds::sincos_cr(k.psi, s, co);
        gv = g.w + g.x * co * dt + g.y * s * dt;
        gpsi = g.z + (g.y * v * co - g.x * v * s) * dt;
        gang = g.z * dt;
    } else {
        float s, co, sb, cb;
        tds::sincos_cr(k.psi + c.ang, s, co);
        tds::sincos_cr(c.ang, sb, cb);
        const float inv_lr = 1.0f / k.lr;
        gv = g.w + g.x * co * dt + g.y * s * dt + g.z * sb * dt * inv_lr;
        const float gth = (g.y * v * co - g.x * v * s) * dt;
        gpsi = g.z + gth;
        gang = gth + g.z * v * cb * dt * inv_lr;
        glr = -g.z * v * sb * dt * inv_lr * inv_lr;
    }
    // v' = v + acc dt; with the no-reversing clamp acc = -v/dt, so d v'/d v = 0 and d v'/d a0 = 0
    float gacc = gv * dt;
    float gv_in = gv;
    if (c.reversing) { gv_in = gv - gacc / dt; gacc = 0.0f; }
    if (p.left_handed) gang = -gang;
    const float ang_scale = k.model == TDS_MODEL_UNICYCLE ? p.max_yaw_rate : p.max_steering;
    if (g_state) reinterpret_cast<float4*>(g_state)[i] = make_float4(g.x, g.y, gpsi, gv_in);
    store_action_grad(g_action, action_dim, i, gacc * p.max_acceleration, gang * ang_scale, 0.0f, 0.0f);
    if (g_lr) g_lr[i] = glr;
}

int check(const float* state, const float* action, int action_dim, const int32_t* model, int64_t n,
          const tds_kinematic_params_t* p, int32_t uniform_model) {
    TDS_REQUIRE(state && action && p, "kinematic: null pointer");
    TDS_REQUIRE(n >= 0, "kinematic: negative n");
    TDS_REQUIRE(action_dim == 2 || action_dim == 4, "kinematic: action_dim must be 2 or 4, got %d", action_dim);
    TDS_REQUIRE(uniform_model >= 0 && uniform_model <= TDS_MODEL_ORIENTED, "kinematic: unknown model %d", uniform_model);
    TDS_REQUIRE(action_dim == 4 || (model == nullptr && uniform_model <= TDS_MODEL_UNICYCLE) || model != nullptr,
                "kinematic: models 3/4 need a 4-value action");
    TDS_REQUIRE(p->dt > 0.0f, "kinematic: dt must be positive");
    return TDS_OK;
}

}  // namespace

extern "C" int tds_kinematic_step_fwd(const float* d_state, const float* d_action, int32_t action_dim, const float* d_lr,
                                      const int32_t* d_model, int32_t uniform_model, int64_t n,
                                      const tds_kinematic_params_t* params, float* d_out_state, void* stream) {
    if (n == 0) return TDS_OK;
    if (int e = check(d_state, d_action, action_dim, d_model, n, params, uniform_model)) return e;
    TDS_REQUIRE(d_out_state, "kinematic: null output");
    if (n == 0) return TDS_OK;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    kin_fwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_state, d_action, action_dim, d_lr, d_model, uniform_model, n,
                                                                *params, d_out_state);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

extern "C" int tds_kinematic_step_bwd(const float* d_state, const float* d_action, int32_t action_dim, const float* d_lr,
                                      const int32_t* d_model, int32_t uniform_model, int64_t n,
                                      const tds_kinematic_params_t* params, const float* d_grad_out,
                                      float* d_grad_state, float* d_grad_action, float* d_grad_lr, void* stream) {
    if (n == 0) return TDS_OK;
    if (int e = check(d_state, d_action, action_dim, d_model, n, params, uniform_model)) return e;
    TDS_REQUIRE(d_grad_out, "kinematic: null grad_out");
    if (n == 0) return TDS_OK;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    kin_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_state, d_action, action_dim, d_lr, d_model, uniform_model, n,
                                                                *params, d_grad_out, d_grad_state, d_grad_action,
                                                                d_grad_lr);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

// ---- replayed NPCs with spawning / despawning: ReplayController.advance_npcs (behavior/replay.py:54-60) followed by
// SpawnController.spawn_despawn_npcs (simulator.py:71-85), one thread per NPC:
//   state, present <- replay[t_replay]                       (if a replay log is given)
//   present &= is_inside_polygon(xy, exit_boundary)          (utils.py:99-122; all edge functions >= 0 or all < 0)
//   spawn = spawn_mask[t_spawn] & !present;  present |= spawn;  state <- spawn ? spawn_state[t_spawn] : state
namespace {
__global__ void __launch_bounds__(128) npc_advance_kernel(const float4* __restrict__ replay, const uint8_t* __restrict__ replay_present,
                                                          int T, int t_replay, const float2* __restrict__ boundary, int V,
                                                          const float4* __restrict__ spawn_states,
                                                          const uint8_t* __restrict__ spawn_masks, int Ts, int t_spawn,
                                                          float4* __restrict__ state, uint8_t* __restrict__ present, int B,
                                                          int Np) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)B * Np) return;
    const int b = (int)(g / Np);
    float4 s = state[g];
    bool here = present[g] != 0;
    if (replay) {
        s = replay[g * T + t_replay];
        here = replay_present ? replay_present[g * T + t_replay] != 0 : true;
    }
    if (boundary) {
        const float2* poly = boundary + (int64_t)b * V;
        int right = 0;
        for (int i = 0; i < V; i++) {
            const float2 p0 = poly[i], p1 = poly[i + 1 == V ? 0 : i + 1];
            const float ea = p1.y - p0.y, eb = p0.x - p1.x;
            const float ec = (-ea) * p0.x - eb * p0.y;
            right += ((ea * s.x + eb * s.y) + ec) >= 0.0f;
        }
        here = here && (right == V || right == 0);
    }
    if (spawn_states) {
        const bool spawn = spawn_masks[g * Ts + t_spawn] != 0 && !here;
        here = here || spawn;
        if (spawn) s = spawn_states[g * Ts + t_spawn];
    }
    state[g] = s;
    present[g] = here ? 1 : 0;
}
}  // namespace

extern "C" int tds_npc_advance(const float* d_replay_states, const uint8_t* d_replay_present, int32_t T, int32_t t_replay,
                               const float* d_exit_boundary, int32_t V, const float* d_spawn_states,
                               const uint8_t* d_spawn_masks, int32_t Ts, int32_t t_spawn, float* d_npc_state,
                               uint8_t* d_npc_present, int32_t B, int32_t Np, void* stream) {
    TDS_REQUIRE(B >= 0 && Np >= 0, "npc_advance: negative size");
    if (B == 0 || Np == 0) return TDS_OK;
    TDS_REQUIRE(d_npc_state && d_npc_present, "npc_advance: null pointer");
    TDS_REQUIRE(!d_replay_states || (T > 0 && t_replay >= 0 && t_replay < T), "npc_advance: replay time %d outside [0,%d)", t_replay, T);
    TDS_REQUIRE(!d_exit_boundary || V >= 1, "npc_advance: empty exit boundary");
    TDS_REQUIRE(!d_spawn_states || (d_spawn_masks && Ts > 0 && t_spawn >= 0 && t_spawn < Ts),
                "npc_advance: spawn time %d outside [0,%d) (or null spawn masks)", t_spawn, Ts);
    const int64_t n = (int64_t)B * Np;
    npc_advance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const float4*)d_replay_states, d_replay_present, T, t_replay, (const float2*)d_exit_boundary, V,
        (const float4*)d_spawn_states, d_spawn_masks, Ts, t_spawn, (float4*)d_npc_state, d_npc_present, B, Np);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

// ---- waypoint goals (goals.py:159-217): collection state[b,a] of M waypoints is achieved when the agent is within
// `threshold` of any of its non-padding waypoints; its mask entries are then cleared and the state advances
// (clamped to the last collection).  One thread per agent.
namespace {
__global__ void __launch_bounds__(128) waypoint_step_kernel(const float4* __restrict__ agent_state, const float2* __restrict__ waypoints,
                                                            uint8_t* __restrict__ mask, int64_t* __restrict__ state,
                                                            int64_t n_agents, int N, int M, float threshold) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_agents) return;
    const int64_t st = min(max(state[g], (int64_t)0), (int64_t)N - 1);
    const int64_t base = (g * N + st) * M;
    const float4 a = agent_state[g];
    bool reached = false;
    for (int m = 0; m < M; m++) {
        const float2 w = waypoints[base + m];
        const float dx = a.x - w.x, dy = a.y - w.y;
        reached |= mask[base + m] != 0 && sqrtf(dx * dx + dy * dy) <= threshold;
    }
    if (reached) {
        for (int m = 0; m < M; m++) mask[base + m] = 0;
        state[g] = min(st + 1, (int64_t)N - 1);
    }
}

// WaypointGoal.get_waypoints / get_masks (goals.py:33-105): the next `count` collections of every agent, zeros past the end
__global__ void __launch_bounds__(256) waypoint_gather_kernel(const float2* __restrict__ waypoints, const uint8_t* __restrict__ mask,
                                                              const int64_t* __restrict__ state, int64_t n_agents, int N, int M,
                                                              int count, float2* __restrict__ out_wp, uint8_t* __restrict__ out_mask) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_agents * count * M) return;
    const int m = (int)(t % M);
    const int c = (int)((t / M) % count);
    const int64_t g = t / ((int64_t)M * count);
    const int64_t idx = state[g] + c;
    const bool valid = idx < N;
    const int64_t src = (g * N + min(max(idx, (int64_t)0), (int64_t)N - 1)) * M + m;
    out_wp[t] = valid ? waypoints[src] : make_float2(0.0f, 0.0f);
    out_mask[t] = valid ? mask[src] : 0;
}
}  // namespace

extern "C" int tds_waypoint_step(const float* d_agent_state, const float* d_waypoints, uint8_t* d_mask, int64_t* d_state,
                                 int64_t n_agents, int32_t N, int32_t M, float threshold, void* stream) {
    TDS_REQUIRE(n_agents >= 0 && N >= 1 && M >= 0, "waypoint_step: need n_agents >= 0, N >= 1, M >= 0");
    if (n_agents == 0 || M == 0) return TDS_OK;
    TDS_REQUIRE(d_agent_state && d_waypoints && d_mask && d_state, "waypoint_step: null pointer");
    waypoint_step_kernel<<<(unsigned)((n_agents + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const float4*)d_agent_state, (const float2*)d_waypoints, d_mask, d_state, n_agents, N, M, threshold);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

extern "C" int tds_waypoint_gather(const float* d_waypoints, const uint8_t* d_mask, const int64_t* d_state, int64_t n_agents,
                                   int32_t N, int32_t M, int32_t count, float* d_out_waypoints, uint8_t* d_out_mask,
                                   void* stream) {
    TDS_REQUIRE(n_agents >= 0 && N >= 1 && M >= 0 && count >= 1, "waypoint_gather: need n_agents >= 0, N >= 1, M >= 0, count >= 1");
    const int64_t n = n_agents * count * M;
    if (n == 0) return TDS_OK;
    TDS_REQUIRE(d_waypoints && d_mask && d_state && d_out_waypoints && d_out_mask, "waypoint_gather: null pointer");
    waypoint_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)d_waypoints, d_mask, d_state, n_agents, N, M, count, (float2*)d_out_waypoints, d_out_mask);
    TDS_LAUNCH_OK();
    return TDS_OK;
}
