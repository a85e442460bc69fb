/* A host with no Python code of its own, driving the model-level C-ABI (include/vf_b200_model.h).  Built and run by
 * tests/test_model_cabi.py:  cabi_host <in.bin> <out.bin>
 * in.bin : int32 n_img, B, T | uint8 images [n_img,32,32,3] | int32 ids [B,T,8,8] | float poses [B,T,7] | uint8 scenes [B,T,32,32,3] | float cameras [B,T,7]
 * out.bin: int64 codes [n_img,8,8] | uint8 decoded [n_img,32,32,3] | int64 codes_last [B,8,8] | int64 query codes [B,8,8] | uint8 generated [B,32,32,3] */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "vf_b200_model.h"

#define CK(x) do { if ((x) != 0) { fprintf(stderr, "%s failed: %s\n", #x, vf_model_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

static const char* VQ_CFG = "{\"ch\": 32, \"ch_mult\": [1, 2, 2], \"attn_resolutions\": [8], \"image_size\": 32, \"embed_dim\": 16, \"z_channels\": 16, \"n_embed\": 64}";
static const char* TR_CFG = "{\"n_layer\": 2, \"n_head\": 4, \"d_model\": 256, \"sequence_size\": 4, \"n_loss_skip\": 1, \"n_embeddings\": 64, \"token_image_size\": 8, \"localization_weight\": \"0\"}";

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 1;
    const int n = hdr[0], B = hdr[1], T = hdr[2];
    const size_t img_b = (size_t)n * 32 * 32 * 3, ids_b = (size_t)B * T * 64 * 4, pose_b = (size_t)B * T * 7 * 4;
    uint8_t* h_img = malloc(img_b); int32_t* h_ids = malloc(ids_b); float* h_pose = malloc(pose_b);
    if (fread(h_img, 1, img_b, f) != img_b || fread(h_ids, 1, ids_b, f) != ids_b || fread(h_pose, 1, pose_b, f) != pose_b) return 1;
    const size_t sc_b = (size_t)B * T * 32 * 32 * 3, cam_b = (size_t)B * T * 7 * 4, gen_b = (size_t)B * 32 * 32 * 3;
    uint8_t* h_sc = malloc(sc_b); float* h_cam = malloc(cam_b); uint8_t* h_gen = malloc(gen_b);
    if (fread(h_sc, 1, sc_b, f) != sc_b || fread(h_cam, 1, cam_b, f) != cam_b) return 1;
    fclose(f);

    vf_handle_t vq = 0, tr = 0, cache = 0;
    CK(vf_vq_create(VQ_CFG, NULL, "fp32", 0, 3, &vq));
    CK(vf_migt_create(TR_CFG, NULL, "fp32", 0, 4, &tr));
    int S, s, K, C, ts, V, mask, loc;
    CK(vf_vq_info(vq, &S, &s, &K, &C));
    CK(vf_migt_info(tr, &ts, &V, &mask, &loc));
    if (S != 32 || s != 8 || K != 64 || C != 3 || ts != 8 || V != 64 || mask != 64) { fprintf(stderr, "unexpected model info\n"); return 4; }

    cudaStream_t st;
    CU(cudaStreamCreate(&st));
    uint8_t *d_img, *d_dec; int64_t *d_codes, *d_last, *d_q; int32_t* d_ids; float* d_pose;
    CU(cudaMalloc((void**)&d_img, img_b)); CU(cudaMalloc((void**)&d_dec, img_b)); CU(cudaMalloc((void**)&d_codes, (size_t)n * 64 * 8));
    CU(cudaMalloc((void**)&d_ids, ids_b)); CU(cudaMalloc((void**)&d_pose, pose_b));
    CU(cudaMalloc((void**)&d_last, (size_t)B * 64 * 8)); CU(cudaMalloc((void**)&d_q, (size_t)B * 64 * 8));
    CU(cudaMemcpyAsync(d_img, h_img, img_b, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ids, h_ids, ids_b, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_pose, h_pose, pose_b, cudaMemcpyHostToDevice, st));

    CK(vf_vq_encode(vq, d_img, VF_LAYOUT_U8_NHWC, n, d_codes, st));
    CK(vf_vq_decode_code(vq, d_codes, n, d_dec, VF_LAYOUT_U8_NHWC, st));
    CK(vf_migt_forward(tr, d_ids, d_pose, B, T, d_last, NULL, st));
    /* context = the first T-1 views; one query per scene at the last view's pose */
    int32_t* h_ctx = malloc((size_t)B * (T - 1) * 64 * 4); float* h_cpose = malloc((size_t)B * (T - 1) * 7 * 4); float* h_qpose = malloc((size_t)B * 7 * 4);
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < (T - 1) * 64; ++i) h_ctx[b * (T - 1) * 64 + i] = h_ids[b * T * 64 + i];
        for (int i = 0; i < (T - 1) * 7; ++i) h_cpose[b * (T - 1) * 7 + i] = h_pose[b * T * 7 + i];
        for (int i = 0; i < 7; ++i) h_qpose[b * 7 + i] = h_pose[b * T * 7 + (T - 1) * 7 + i];
    }
    int32_t* d_ctx; float *d_cpose, *d_qpose;
    CU(cudaMalloc((void**)&d_ctx, (size_t)B * (T - 1) * 64 * 4)); CU(cudaMalloc((void**)&d_cpose, (size_t)B * (T - 1) * 7 * 4)); CU(cudaMalloc((void**)&d_qpose, (size_t)B * 7 * 4));
    CU(cudaMemcpyAsync(d_ctx, h_ctx, (size_t)B * (T - 1) * 64 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_cpose, h_cpose, (size_t)B * (T - 1) * 7 * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_qpose, h_qpose, (size_t)B * 7 * 4, cudaMemcpyHostToDevice, st));
    CK(vf_migt_prefill_context(tr, d_ctx, d_cpose, B, T - 1, &cache, st));
    CK(vf_migt_query(tr, cache, d_qpose, B, d_q, st));
    /* the whole of generate_batch_predictions: uint8 scenes + world cameras in, the novel view out */
    uint8_t *d_sc, *d_gen; float* d_cam;
    CU(cudaMalloc((void**)&d_sc, sc_b)); CU(cudaMalloc((void**)&d_cam, cam_b)); CU(cudaMalloc((void**)&d_gen, gen_b));
    CU(cudaMemcpyAsync(d_sc, h_sc, sc_b, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_cam, h_cam, cam_b, cudaMemcpyHostToDevice, st));
    CK(vf_generate(tr, vq, d_sc, d_cam, B, T, d_gen, NULL, st));
    CU(cudaMemcpyAsync(h_gen, d_gen, gen_b, cudaMemcpyDeviceToHost, st));

    int64_t* h_codes = malloc((size_t)n * 64 * 8); uint8_t* h_dec = malloc(img_b); int64_t* h_last = malloc((size_t)B * 64 * 8); int64_t* h_q = malloc((size_t)B * 64 * 8);
    CU(cudaMemcpyAsync(h_codes, d_codes, (size_t)n * 64 * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_dec, d_dec, img_b, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_last, d_last, (size_t)B * 64 * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_q, d_q, (size_t)B * 64 * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    /* error path: a bad handle must come back as a status + message, not a crash */
    if (vf_vq_encode(9999, d_img, 0, n, d_codes, st) == 0) { fprintf(stderr, "bad handle accepted\n"); return 5; }
    CK(vf_destroy(cache)); CK(vf_destroy(tr)); CK(vf_destroy(vq));

    f = fopen(argv[2], "wb");
    fwrite(h_codes, 8, (size_t)n * 64, f); fwrite(h_dec, 1, img_b, f); fwrite(h_last, 8, (size_t)B * 64, f); fwrite(h_q, 8, (size_t)B * 64, f); fwrite(h_gen, 1, gen_b, f);
    fclose(f);
    printf("cabi_host ok: %d images, %d scenes x %d views; last error after the bad-handle probe: %s\n", n, B, T, vf_model_last_error());
    return 0;
}
