// Graph-captured DDIM sampling of one generated frame: stands in for the inner loop of reference
// generate.py:204-220 + train_dit.py:30-125 (101 x [build t / t_next, window slice, DiT forward,
// v -> x0 -> eps -> DDIM update, write back the last frame]).
//
// One CUDA graph holds a whole step: step_prep (timestep rows + DDIM coefficients from a device-side
// step counter) -> DiT backbone on the window -> fused DDIM update of the last frame.  The host
// replays it noise_steps+1 times per frame with no per-step host data, no sync and no allocation.
// The conditioning table for the frame (context rows at the stabilisation level, one row per noise
// level for the last frame) is computed once per frame by gtav_dit_conditioning: adaLN depends only
// on (t, action), not on x, so this is an exact hoist.
#include <new>

#include "../../include/gtav_b200.h"
#include "kernels.h"

using namespace gtav;

struct gtav_sampler_s {
    gtav_dit_plan_t plan;
    int B, T, steps, n;            // n = elements per latent frame
    float* x_win;                  // fp32 [B, T, n] window state (last frame is the one being denoised)
    bf16* v_out;                   // bf16 [B, T, n]
    const float* abar;             // device [max_noise_level]
    // scratch (device)
    int* counter; int* levels; int* frame_row; int* final_flag; float* abar_t; float* abar_next;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int use_graph;
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
    s->abar_t = static_cast<float*>(take(sizeof(float) * s->B));
    s->abar_next = static_cast<float*>(take(sizeof(float) * s->B));
    return off;
}

int enqueue_step(gtav_sampler_s* s, cudaStream_t st) {
    int rc = launch_step_prep(s->counter, s->levels, s->abar, s->B, s->T, s->steps, s->frame_row, s->abar_t, s->abar_next,
                              s->final_flag, st);
    if (rc) return rc;
    if ((rc = gtav_dit_backbone(s->plan, s->x_win, 0, s->frame_row, s->v_out, st))) return rc;
    const long fs = static_cast<long>(s->T) * s->n;
    const long last = static_cast<long>(s->T - 1) * s->n;
    return launch_ddim(s->x_win + last, fs, s->v_out + last, fs, s->x_win + last, fs, s->B, s->n, s->abar_t, s->abar_next,
                       s->final_flag, st);
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
                        const float* abar_dev, const int* levels_host, void* scratch, size_t scratch_bytes, int use_graph,
                        gtav_stream_t stream, gtav_sampler_t* out) {
    if (!plan || !x_win || !v_out || !abar_dev || !levels_host || !scratch || !out || B <= 0 || T <= 0 || steps < 0) {
        set_error("sampler_create: bad argument");
        return -1;
    }
    gtav_sampler_s* s = new (std::nothrow) gtav_sampler_s();
    if (!s) { set_error("sampler_create: out of host memory"); return -4; }
    s->plan = plan; s->B = B; s->T = T; s->steps = steps; s->n = frame_elems;
    s->x_win = x_win; s->v_out = static_cast<bf16*>(v_out); s->abar = abar_dev;
    s->use_graph = use_graph;
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
    if (s->exec) cudaGraphExecDestroy(s->exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    delete s;
}

int gtav_sampler_run_frame(gtav_sampler_t s, int n_steps, gtav_stream_t stream) {
    if (!s) { set_error("sampler_run_frame: null handle"); return -1; }
    if (n_steps < 0 || n_steps > s->steps + 1) n_steps = s->steps + 1;
    int rc0 = launch_set_int(s->counter, s->steps, stream);      // k counts down from `steps` to 0
    if (rc0) return rc0;
    if (n_steps == 0) return 0;
    if (!s->use_graph) {
        for (int i = 0; i < n_steps; ++i) {
            int rc = enqueue_step(s, stream);
            if (rc) return rc;
        }
        return 0;
    }
    if (!s->exec) {
        // warm-up outside capture (first-use cudaFuncSetAttribute calls, lazy module load), then restore the counter
        int rc = enqueue_step(s, stream);
        if (rc) return rc;
        GTAV_CUDA_OK(cudaStreamSynchronize(stream));
        GTAV_CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_step(s, stream);
        cudaError_t e = cudaStreamEndCapture(stream, &s->graph);
        if (rc) return rc;
        if (e != cudaSuccess) { set_error("sampler: graph capture failed: %s", cudaGetErrorString(e)); return -2; }
        GTAV_CUDA_OK(cudaGraphInstantiate(&s->exec, s->graph, 0));
        // the warm-up step consumed one real step (k = steps): continue from there
        for (int i = 1; i < n_steps; ++i) GTAV_CUDA_OK(cudaGraphLaunch(s->exec, stream));
        return 0;
    }
    for (int i = 0; i < n_steps; ++i) GTAV_CUDA_OK(cudaGraphLaunch(s->exec, stream));
    return 0;
}

int gtav_noise_clamp(const float* noise, float* x, long x_stride, int F, int n, float amax, gtav_stream_t stream) {
    return launch_noise_clamp(noise, x, x_stride, F, n, amax, stream);
}

}  // extern "C"
