// Graph-captured DDIM sampling of one generated frame: stands in for the inner loop of reference
// generate.py:204-220 + train_dit.py:30-125 (101 x [build t / t_next, window slice, DiT forward,
// v -> x0 -> eps -> DDIM update, write back the last frame]).
//
// A step = step_prep (timestep rows + DDIM coefficients from a device-side step counter) -> DiT backbone ->
// fused DDIM update of the last frame, all enqueued with programmatic dependent launch; the whole frame
// (noise_steps + 1 steps) is captured into ONE CUDA graph and replayed per generated frame, so there is no
// per-step host work, sync or allocation.
//
// Two exact algorithmic hoists, both parity-tested against the dense path (tests/test_sampler_gpu.py):
//   * conditioning table: adaLN depends only on (t, action), not on x - computed
//     once per frame by gtav_dit_conditioning for all noise levels;
//   * frame cache (GTAV_SAMPLER_FRAME_CACHE): attention is per-frame (spatial) or causal over frames (temporal)
//     and the context frames, their timestep and actions do not change during the 101 steps of a frame
//     (SURVEY.md section 0, fact 1), so their activations are identical at every step: one context pass stores
//     the rotated K and V of every temporal layer, and each step recomputes only the frame being denoised
//     (M = 144*B rows instead of 720*B) against that cache.
#include <new>

#include "../../include/gtav_b200.h"
#include "kernels.h"

using namespace gtav;

struct gtav_sampler_s {
    gtav_dit_plan_t plan;
    int B, T, steps, n;            // n = elements per latent frame
    float* x_win;                  // fp32 [B, T, n] window state (last frame is the one being denoised)
    bf16* v_out;                   // bf16 [B, T, n] (dense) / [B, n] in its first rows (frame cache)
    const float* abar;             // device [max_noise_level]
    // scratch (device)
    int* counter; int* levels; int* frame_row; int* last_row; int* final_flag; float* abar_t; float* abar_next;
    cudaGraphExec_t exec_step;     // one step
    cudaGraphExec_t exec_frame;    // context pass (if cached) + all steps+1 steps
    int use_graph, use_cache;
};

namespace {
size_t scratch_layout(gtav_sampler_s* s, void* base) {
    uint8_t* p = static_cast<uint8_t*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* r = p + off; off += (bytes + 255) & ~size_t(255); return r; };
    s->counter = static_cast<int*>(take(sizeof(int)));
    s->final_flag = static_cast<int*>(take(sizeof(int)));
    s->levels = static_cast<int*>(take(sizeof(int) * (s->steps + 1)));
    s->frame_row = static_cast<int*>(take(sizeof(int) * s->B * s->T));
    s->last_row = static_cast<int*>(take(sizeof(int) * s->B));
    s->abar_t = static_cast<float*>(take(sizeof(float) * s->B));
    s->abar_next = static_cast<float*>(take(sizeof(float) * s->B));
    return off;
}

int enqueue_context(gtav_sampler_s* s, cudaStream_t st) {
    if (!s->use_cache) return 0;
    return gtav_dit_context(s->plan, s->x_win, 0, nullptr, st);      // context rows of the table are rows 0 .. B*(T-1)-1
}

int enqueue_step(gtav_sampler_s* s, cudaStream_t st) {
    int rc = launch_step_prep(s->counter, s->levels, s->abar, s->B, s->T, s->steps, s->frame_row, s->last_row, s->abar_t,
                              s->abar_next, s->final_flag, st);
    if (rc) return rc;
    const long fs = static_cast<long>(s->T) * s->n;
    const long last = static_cast<long>(s->T - 1) * s->n;
    if (s->use_cache) {
        if ((rc = gtav_dit_last_frame(s->plan, s->x_win, 0, s->last_row, s->v_out, st))) return rc;
        return launch_ddim(s->x_win + last, fs, s->v_out, s->n, s->x_win + last, fs, s->B, s->n, s->abar_t, s->abar_next,
                           s->final_flag, st);
    }
    if ((rc = gtav_dit_backbone(s->plan, s->x_win, 0, s->frame_row, s->v_out, st))) return rc;
    return launch_ddim(s->x_win + last, fs, s->v_out + last, fs, s->x_win + last, fs, s->B, s->n, s->abar_t, s->abar_next,
                       s->final_flag, st);
}

// Capture fn(stream) into an executable graph.
template <typename Fn>
int capture(cudaStream_t stream, cudaGraphExec_t* exec, Fn fn) {
    cudaGraph_t graph = nullptr;
    GTAV_CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    const int rc = fn();
    const cudaError_t e = cudaStreamEndCapture(stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { set_error("sampler: graph capture failed: %s", cudaGetErrorString(e)); return -2; }
    const cudaError_t e2 = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { set_error("sampler: graph instantiation failed: %s", cudaGetErrorString(e2)); return -2; }
    return 0;
}
}  // namespace

extern "C" {

size_t gtav_sampler_scratch_bytes(int B, int T, int steps) {
    gtav_sampler_s tmp{};
    tmp.B = B; tmp.T = T; tmp.steps = steps;
    return scratch_layout(&tmp, nullptr);
}

int gtav_sampler_cond_rows(int B, int T, int steps) { return B * (T - 1) + B * (steps + 1); }

int gtav_sampler_create(gtav_dit_plan_t plan, int B, int T, int steps, int frame_elems, float* x_win, void* v_out,
                        const float* abar_dev, const int* levels_host, void* scratch, size_t scratch_bytes, int flags,
                        gtav_stream_t stream, gtav_sampler_t* out) {
    if (!plan || !x_win || !v_out || !abar_dev || !levels_host || !scratch || !out || B <= 0 || T <= 0 || steps < 0) {
        set_error("sampler_create: bad argument");
        return -1;
    }
    gtav_sampler_s* s = new (std::nothrow) gtav_sampler_s();
    if (!s) { set_error("sampler_create: out of host memory"); return -4; }
    s->plan = plan; s->B = B; s->T = T; s->steps = steps; s->n = frame_elems;
    s->x_win = x_win; s->v_out = static_cast<bf16*>(v_out); s->abar = abar_dev;
    s->use_graph = (flags & GTAV_SAMPLER_GRAPH) != 0;
    s->use_cache = (flags & GTAV_SAMPLER_FRAME_CACHE) != 0;
    if (scratch_layout(s, scratch) > scratch_bytes) {
        set_error("sampler_create: scratch too small");
        delete s;
        return -1;
    }
    cudaError_t e = cudaMemcpyAsync(s->levels, levels_host, sizeof(int) * (steps + 1), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);      // levels_host may be a temporary
    if (e != cudaSuccess) {
        set_error("sampler_create: copying the noise levels failed: %s", cudaGetErrorString(e));
        delete s;
        return -2;
    }
    *out = s;
    return 0;
}

void gtav_sampler_destroy(gtav_sampler_t s) {
    if (!s) return;
    if (s->exec_step) cudaGraphExecDestroy(s->exec_step);
    if (s->exec_frame) cudaGraphExecDestroy(s->exec_frame);
    delete s;
}

int gtav_sampler_run_frame(gtav_sampler_t s, int n_steps, gtav_stream_t stream) {
    if (!s) { set_error("sampler_run_frame: null handle"); return -1; }
    if (n_steps < 0 || n_steps > s->steps + 1) n_steps = s->steps + 1;
    int rc = launch_set_int(s->counter, s->steps, stream);      // k counts down from `steps` to 0
    if (rc) return rc;
    if (n_steps == 0) return 0;
    if (!s->use_graph) {
        if ((rc = enqueue_context(s, stream))) return rc;
        for (int i = 0; i < n_steps; ++i)
            if ((rc = enqueue_step(s, stream))) return rc;
        return 0;
    }
    const bool whole = n_steps == s->steps + 1;
    if ((whole && !s->exec_frame) || (!whole && !s->exec_step)) {
        // first use: run the frame eagerly once (first-use cudaFuncSetAttribute calls and lazy module loading are
        // not capturable), then capture for the following frames
        if ((rc = enqueue_context(s, stream))) return rc;
        for (int i = 0; i < n_steps; ++i)
            if ((rc = enqueue_step(s, stream))) return rc;
        GTAV_CUDA_OK(cudaStreamSynchronize(stream));
        if (whole) {
            return capture(stream, &s->exec_frame, [&]() {
                int r = enqueue_context(s, stream);
                for (int i = 0; i < n_steps && r == 0; ++i) r = enqueue_step(s, stream);
                return r;
            });
        }
        return capture(stream, &s->exec_step, [&]() { return enqueue_step(s, stream); });
    }
    if (whole) {
        GTAV_CUDA_OK(cudaGraphLaunch(s->exec_frame, stream));
        return 0;
    }
    if ((rc = enqueue_context(s, stream))) return rc;
    for (int i = 0; i < n_steps; ++i) GTAV_CUDA_OK(cudaGraphLaunch(s->exec_step, stream));
    return 0;
}

int gtav_noise_clamp(const float* noise, float* x, long x_stride, int F, int n, float amax, gtav_stream_t stream) {
    return launch_noise_clamp(noise, x, x_stride, F, n, amax, stream);
}

}  // extern "C"
