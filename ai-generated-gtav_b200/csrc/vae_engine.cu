// Host-side VAE engine: launch plans for AutoencoderKL.encode / decode of the ViT-L-20 VAE
// (reference model/vae.py:306-338 with AttentionBlock 154-157 and Attention 78-112).
#include <new>
#include <vector>

#include "../../include/gtav_b200.h"
#include "kernels.h"

using namespace gtav;

struct gtav_vae_s {
    gtav_vae_config cfg;
    gtav_vae_weights w;
    std::vector<gtav_vae_block> enc, dec;
    int seq, patch_dim, patch_ld;   // 576, 1200, 1200
};

struct BlockOps { GemmOp qkv, proj, fc1, fc2; };

struct gtav_vae_plan_s {
    gtav_vae_t eng;
    int N, M;
    bf16 *pa, *h, *hn, *qkv, *att, *mlp, *mom, *zin, *pred;
    GemmOp g_patch, g_quant, g_post, g_pred;
    std::vector<BlockOps> enc, dec;
};

namespace {

struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
    bf16* take(size_t elems) {
        bf16* r = reinterpret_cast<bf16*>(base + off);
        off += (elems * sizeof(bf16) + 1023) & ~size_t(1023);
        return r;
    }
};

void carve(gtav_vae_plan_s* p, void* ws, size_t* total) {
    const gtav_vae_s* e = p->eng;
    const size_t M = p->M, D = e->cfg.dim;
    Carver c(ws);
    p->pa = c.take(M * e->patch_ld);     // patchified image; reused as the predictor output
    p->h = c.take(M * D);
    p->hn = c.take(M * D);
    p->qkv = c.take(M * 3 * D);
    p->att = c.take(M * D);
    p->mlp = c.take(M * 4 * D);
    p->mom = c.take(M * 64);
    p->zin = c.take(M * 64);
    p->pred = p->pa;
    *total = c.off;
}

GemmParams gp(bf16* out, int ldo, const void* bias, int M, int N, int K) {
    GemmParams p{};
    p.out = out; p.ldo = ldo; p.bias = static_cast<const bf16*>(bias);
    p.M = M; p.N = N; p.K = K; p.rows_per_frame = 1;
    return p;
}

int prepare_block(gtav_vae_plan_s* p, const gtav_vae_block& b, BlockOps* o) {
    const int D = p->eng->cfg.dim, M = p->M;
    int rc = gemm_prepare(&o->qkv, p->hn, D, static_cast<const bf16*>(b.qkv_w), D, gp(p->qkv, 3 * D, b.qkv_b, M, 3 * D, D), EPI_BIAS);
    GemmParams q = gp(p->h, D, b.proj_b, M, D, D);
    q.res = p->h; q.ldr = D;
    rc |= gemm_prepare(&o->proj, p->att, D, static_cast<const bf16*>(b.proj_w), D, q, EPI_BIAS_RES);
    rc |= gemm_prepare(&o->fc1, p->hn, D, static_cast<const bf16*>(b.fc1_w), D, gp(p->mlp, 4 * D, b.fc1_b, M, 4 * D, D), EPI_BIAS_GELU_ERF);
    GemmParams r = gp(p->h, D, b.fc2_b, M, D, 4 * D);
    r.res = p->h; r.ldr = D;
    rc |= gemm_prepare(&o->fc2, p->mlp, 4 * D, static_cast<const bf16*>(b.fc2_w), 4 * D, r, EPI_BIAS_RES);
    return rc;
}

int run_block(gtav_vae_plan_s* p, const gtav_vae_block& b, const BlockOps& o, cudaStream_t s) {
    const gtav_vae_s* e = p->eng;
    const int D = e->cfg.dim;
    int rc;
    if ((rc = launch_ln_affine(p->h, p->hn, p->M, D, b.norm1_w, b.norm1_b, s))) return rc;
    if ((rc = gemm_run(&o.qkv, s))) return rc;
    if ((rc = launch_attention_seq(p->qkv, p->att, p->N, e->seq, e->cfg.heads, reinterpret_cast<const float2*>(e->w.rot), 16, s))) return rc;
    if ((rc = gemm_run(&o.proj, s))) return rc;
    if ((rc = launch_ln_affine(p->h, p->hn, p->M, D, b.norm2_w, b.norm2_b, s))) return rc;
    if ((rc = gemm_run(&o.fc1, s))) return rc;
    return gemm_run(&o.fc2, s);
}

}  // namespace

extern "C" {

int gtav_vae_create(const gtav_vae_config* cfg, const gtav_vae_weights* w, gtav_vae_t* out) {
    if (!cfg || !w || !out) { set_error("vae_create: null argument"); return -1; }
    if (cfg->dim != 1024 || cfg->heads != 16 || cfg->seq_h * cfg->seq_w != 576 || cfg->latent_dim * 2 > 64 ||
        (3 * cfg->patch * cfg->patch) % 8 != 0 || cfg->enc_depth < 0 || cfg->dec_depth < 0) {
        set_error("vae_create: unsupported geometry (dim=%d heads=%d seq=%dx%d latent=%d patch=%d); kernels are built for "
                  "dim 1024, 16 heads of 64, 576 tokens", cfg->dim, cfg->heads, cfg->seq_h, cfg->seq_w, cfg->latent_dim, cfg->patch);
        return -1;
    }
    gtav_vae_s* e = new (std::nothrow) gtav_vae_s();
    if (!e) { set_error("vae_create: out of host memory"); return -4; }
    e->cfg = *cfg;
    e->w = *w;
    e->enc.assign(w->enc, w->enc + cfg->enc_depth);
    e->dec.assign(w->dec, w->dec + cfg->dec_depth);
    e->seq = cfg->seq_h * cfg->seq_w;
    e->patch_dim = 3 * cfg->patch * cfg->patch;
    e->patch_ld = e->patch_dim;
    *out = e;
    return 0;
}

void gtav_vae_destroy(gtav_vae_t h) { delete h; }

size_t gtav_vae_workspace_bytes(gtav_vae_t h, int n_frames) {
    if (!h || n_frames <= 0) return 0;
    gtav_vae_plan_s tmp{};
    tmp.eng = h; tmp.N = n_frames; tmp.M = n_frames * h->seq;
    size_t total = 0;
    carve(&tmp, nullptr, &total);
    return total;
}

int gtav_vae_plan_create(gtav_vae_t h, int n_frames, void* workspace, size_t workspace_bytes, gtav_vae_plan_t* out) {
    if (!h || !workspace || !out || n_frames <= 0) { set_error("vae_plan_create: bad argument"); return -1; }
    if (reinterpret_cast<uintptr_t>(workspace) & 1023) { set_error("vae_plan_create: workspace must be 1024-byte aligned"); return -1; }
    gtav_vae_plan_s* p = new (std::nothrow) gtav_vae_plan_s();
    if (!p) { set_error("vae_plan_create: out of host memory"); return -4; }
    p->eng = h; p->N = n_frames; p->M = n_frames * h->seq;
    size_t need = 0;
    carve(p, workspace, &need);
    if (need > workspace_bytes) {
        set_error("vae_plan_create: workspace too small (%zu < %zu)", workspace_bytes, need);
        delete p;
        return -1;
    }
    const int D = h->cfg.dim, M = p->M, L = h->cfg.latent_dim;
    const gtav_vae_weights& w = h->w;
    int rc = 0;
    rc |= gemm_prepare(&p->g_patch, p->pa, h->patch_ld, static_cast<const bf16*>(w.patch_w), h->patch_ld,
                       gp(p->h, D, w.patch_b, M, D, h->patch_dim), EPI_BIAS);
    rc |= gemm_prepare(&p->g_quant, p->hn, D, static_cast<const bf16*>(w.quant_w), D, gp(p->mom, 64, w.quant_b, M, 2 * L, D), EPI_BIAS);
    rc |= gemm_prepare(&p->g_post, p->zin, 64, static_cast<const bf16*>(w.post_w), 64, gp(p->h, D, w.post_b, M, D, 64), EPI_BIAS);
    rc |= gemm_prepare(&p->g_pred, p->hn, D, static_cast<const bf16*>(w.pred_w), D,
                       gp(p->pred, h->patch_ld, w.pred_b, M, h->patch_dim, D), EPI_BIAS);
    p->enc.resize(h->enc.size());
    p->dec.resize(h->dec.size());
    for (size_t i = 0; i < h->enc.size() && rc == 0; ++i) rc |= prepare_block(p, h->enc[i], &p->enc[i]);
    for (size_t i = 0; i < h->dec.size() && rc == 0; ++i) rc |= prepare_block(p, h->dec[i], &p->dec[i]);
    if (rc) { delete p; return rc < 0 ? rc : -1; }
    *out = p;
    return 0;
}

void gtav_vae_plan_destroy(gtav_vae_plan_t p) { delete p; }

static int vae_encode_moments(gtav_vae_plan_t p, const void* img, int img_is_bf16, gtav_stream_t stream);

int gtav_vae_encode(gtav_vae_plan_t p, const void* img, int img_is_bf16, float* mean_out, float scale, int round_bf16,
                    gtav_stream_t stream) {
    if (!p || !img || !mean_out) { set_error("vae_encode: null argument"); return -1; }
    int rc = vae_encode_moments(p, img, img_is_bf16, stream);
    if (rc) return rc;
    return launch_take_mean(p->mom, 64, mean_out, p->M, p->eng->cfg.latent_dim, scale, round_bf16, stream);
}

int gtav_vae_encode_moments(gtav_vae_plan_t p, const void* img, int img_is_bf16, float* mean_out, float* logvar_out,
                            gtav_stream_t stream) {
    if (!p || !img || !mean_out || !logvar_out) { set_error("vae_encode_moments: null argument"); return -1; }
    int rc = vae_encode_moments(p, img, img_is_bf16, stream);
    if (rc) return rc;
    const int C = p->eng->cfg.latent_dim;
    if ((rc = launch_take_mean(p->mom, 64, mean_out, p->M, C, 1.0f, 0, stream))) return rc;
    return launch_take_mean(p->mom + C, 64, logvar_out, p->M, C, 1.0f, 0, stream);     // columns C .. 2C-1 of quant_conv
}

// patchify -> patch embed -> encoder blocks -> enc_norm -> quant_conv: the 2*latent moments of every token in p->mom
static int vae_encode_moments(gtav_vae_plan_t p, const void* img, int img_is_bf16, gtav_stream_t stream) {
    const gtav_vae_s* e = p->eng;
    const gtav_vae_config& c = e->cfg;
    int rc = launch_patchify(img, img_is_bf16, p->pa, e->patch_ld, p->N, 3, c.seq_h * c.patch, c.seq_w * c.patch, c.patch, 0, 0L, stream);
    if (rc) return rc;
    if ((rc = gemm_run(&p->g_patch, stream))) return rc;
    for (size_t i = 0; i < e->enc.size(); ++i)
        if ((rc = run_block(p, e->enc[i], p->enc[i], stream))) return rc;
    if ((rc = launch_ln_affine(p->h, p->hn, p->M, c.dim, e->w.enc_norm_w, e->w.enc_norm_b, stream))) return rc;
    return gemm_run(&p->g_quant, stream);
}

int gtav_vae_decode(gtav_vae_plan_t p, const float* z, float divisor, void* out, int to_u8, gtav_stream_t stream) {
    if (!p || !z || !out) { set_error("vae_decode: null argument"); return -1; }
    const gtav_vae_s* e = p->eng;
    const gtav_vae_config& c = e->cfg;
    int rc = launch_cast_pad(z, p->zin, p->M, c.latent_dim, 64, divisor, stream);
    if (rc) return rc;
    if ((rc = gemm_run(&p->g_post, stream))) return rc;
    for (size_t i = 0; i < e->dec.size(); ++i)
        if ((rc = run_block(p, e->dec[i], p->dec[i], stream))) return rc;
    if ((rc = launch_ln_affine(p->h, p->hn, p->M, c.dim, e->w.dec_norm_w, e->w.dec_norm_b, stream))) return rc;
    if ((rc = gemm_run(&p->g_pred, stream))) return rc;
    return launch_vae_unpatchify(p->pred, out, to_u8, p->N, c.seq_h, c.seq_w, c.patch, stream);
}

}  // extern "C"
