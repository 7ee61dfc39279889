// Glue kernels of the per-step path and of differentiable rollouts: the small tensor plumbing the reference does
// with eager torch ops between its hot functions, as single launches of our own.
//
//   tds_agent_boxes      state (x, y, psi, v) + size (l, w) -> collision boxes (x, y, l, w, psi)        simulator.py:1161-1170
//                        (+ the position and the (sin, cos) of the heading of every agent = its egocentric camera,   simulator.py:961, 1017)
//   tds_rollout_loss     acc += [sum collision, sum offroad, sum |xy - target|^2] of one step             imitation_learning.py:279-335
//   tds_rollout_grad     d loss / d state of one step: the gradient that arrives from the next kinematic step
//                        + offroad backward + collision backward (scattered from box layout) + the target term
//
// A differentiable rollout (BASELINE config 5) is then a fixed sequence of launches - kinematic forward, boxes,
// collisions, offroad, loss; and backwards: offroad backward, collision backward, this merge, kinematic backward - that
// the host layer records once into a CUDA graph (torchdrivesim_b200/rollout.py) instead of driving ~30 small
// launches per step through the autograd tape.
#include "tds_common.cuh"

namespace {

__global__ void __launch_bounds__(256) agent_boxes_kernel(const float4* __restrict__ state, const float2* __restrict__ size, int64_t n,
                                                          float* __restrict__ box, float2* __restrict__ cam_sc, float2* __restrict__ xy) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = state[i];
    if (box) {
        const float2 z = size[i];
        float* o = box + 5 * i;
        o[0] = s.x; o[1] = s.y; o[2] = z.x; o[3] = z.y; o[4] = s.z;
    }
    if (cam_sc) {
        float sn, cs;
        tds::sincos_cr(s.z, sn, cs);
        cam_sc[i] = make_float2(sn, cs);
    }
    if (xy) xy[i] = make_float2(s.x, s.y);
}

// ONE CTA, fixed reduction order (reproducible sums), like tds_infraction_metrics
__global__ void __launch_bounds__(1024) rollout_loss_kernel(const float* __restrict__ collision, const float* __restrict__ offroad,
                                                            const float4* __restrict__ state, const float2* __restrict__ target,
                                                            int64_t n, double* __restrict__ acc) {
    __shared__ double s_part[32][3];
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (collision) v[0] += (double)collision[i];
        if (offroad) v[1] += (double)offroad[i];
        if (target) {
            const float4 s = state[i];
            const float2 t = target[i];
            const float dx = s.x - t.x, dy = s.y - t.y;
            v[2] += (double)(dx * dx) + (double)(dy * dy);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 3; k++) s_part[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w][threadIdx.x];
        acc[threadIdx.x] += t;
    }
}

__global__ void __launch_bounds__(256) rollout_grad_kernel(const float4* __restrict__ g_next, const float4* __restrict__ g_offroad,
                                                           const float* __restrict__ g_box_ego, const float* __restrict__ g_box_all,
                                                           const float4* __restrict__ state, const float2* __restrict__ target,
                                                           float w_offroad, float w_collision, float w_target, int64_t n,
                                                           float4* __restrict__ g_state) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 g = g_next ? g_next[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (g_offroad) {
        const float4 o = g_offroad[i];
        g.x += w_offroad * o.x; g.y += w_offroad * o.y; g.z += w_offroad * o.z;
    }
    if (g_box_ego) {
        const float* e = g_box_ego + 5 * i;
        float bx = e[0], by = e[1], bp = e[4];
        if (g_box_all) {
            const float* a = g_box_all + 5 * i;
            bx += a[0]; by += a[1]; bp += a[4];
        }
        g.x += w_collision * bx; g.y += w_collision * by; g.z += w_collision * bp;
    }
    if (target) {
        const float4 s = state[i];
        const float2 t = target[i];
        g.x += w_target * (s.x - t.x);
        g.y += w_target * (s.y - t.y);
    }
    g_state[i] = g;
}

}  // namespace

extern "C" int tds_agent_boxes(const float* d_state, const float* d_size, int64_t n, float* d_box, float* d_cam_sc, float* d_xy,
                               void* stream) {
    TDS_REQUIRE(n >= 0, "agent_boxes: negative n");
    if (n == 0) return TDS_OK;
    TDS_REQUIRE(d_state && (d_box == nullptr || d_size) && (d_box || d_cam_sc || d_xy), "agent_boxes: null pointer");
    agent_boxes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_state), reinterpret_cast<const float2*>(d_size), n, d_box, reinterpret_cast<float2*>(d_cam_sc),
        reinterpret_cast<float2*>(d_xy));
    TDS_LAUNCH_OK();
    return TDS_OK;
}

extern "C" int tds_rollout_loss(const float* d_collision, const float* d_offroad, const float* d_state, const float* d_target_xy,
                                int64_t n, double* d_acc, void* stream) {
    TDS_REQUIRE(n >= 0 && d_acc, "rollout_loss: bad arguments");
    TDS_REQUIRE(d_target_xy == nullptr || d_state, "rollout_loss: the target term needs the state");
    if (n == 0) return TDS_OK;
    rollout_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_collision, d_offroad, reinterpret_cast<const float4*>(d_state),
                                                             reinterpret_cast<const float2*>(d_target_xy), n, d_acc);
    TDS_LAUNCH_OK();
    return TDS_OK;
}

extern "C" int tds_rollout_grad(const float* d_grad_next, const float* d_grad_offroad, const float* d_grad_box_ego,
                                const float* d_grad_box_all, const float* d_state, const float* d_target_xy, float w_offroad,
                                float w_collision, float w_target, int64_t n, float* d_grad_state, void* stream) {
    TDS_REQUIRE(n >= 0 && d_grad_state, "rollout_grad: bad arguments");
    TDS_REQUIRE(d_target_xy == nullptr || d_state, "rollout_grad: the target term needs the state");
    TDS_REQUIRE(d_grad_box_all == nullptr || d_grad_box_ego, "rollout_grad: grad_box_all without grad_box_ego");
    if (n == 0) return TDS_OK;
    rollout_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_grad_next), reinterpret_cast<const float4*>(d_grad_offroad), d_grad_box_ego, d_grad_box_all,
        reinterpret_cast<const float4*>(d_state), reinterpret_cast<const float2*>(d_target_xy), w_offroad, w_collision, w_target, n,
        reinterpret_cast<float4*>(d_grad_state));
    TDS_LAUNCH_OK();
    return TDS_OK;
}
