// jpeg_io.cuh -- device-side JPEG decode / encode around the hot path: the GPU counterparts of
// cv2.imread (reference app.py:122,154,158; ransac.py:291-294) and cv2.imwrite (app.py:193-195; ransac.py:301), so that a
// frame can go file bytes -> HBM -> warp / stitch -> HBM -> file bytes without a decoded copy ever crossing the host link.
//
// The codec is NVIDIA's nvJPEG (library code, like cuBLAS for a plain GEMM: entropy decoding + IDCT + colour conversion
// are not this framework's path).  It is loaded lazily with dlopen so that libb200geo.so carries no link-time dependency
// on it: every other entry point works on a box without libnvjpeg.  What is ours here is the glue the reference's callers
// need: interleaved BGR (cv2's order) or RGB (the order ransac.example1 / app.py flip to), pitch-aware device buffers, and
// cv2.imwrite's defaults (quality 95, 4:2:0 chroma, baseline Huffman).
#pragma once

#include "common.cuh"

#include <dlfcn.h>
#include <nvjpeg.h>

namespace b200geo {

struct JpegApi {
    void *lib;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t *);
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t);
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t *);
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t);
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char *, size_t, int *, nvjpegChromaSubsampling_t *, int *, int *);
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *, size_t, nvjpegOutputFormat_t, nvjpegImage_t *, cudaStream_t);
    nvjpegStatus_t (*EncoderStateCreate)(nvjpegHandle_t, nvjpegEncoderState_t *, cudaStream_t);
    nvjpegStatus_t (*EncoderStateDestroy)(nvjpegEncoderState_t);
    nvjpegStatus_t (*EncoderParamsCreate)(nvjpegHandle_t, nvjpegEncoderParams_t *, cudaStream_t);
    nvjpegStatus_t (*EncoderParamsDestroy)(nvjpegEncoderParams_t);
    nvjpegStatus_t (*EncoderParamsSetQuality)(nvjpegEncoderParams_t, const int, cudaStream_t);
    nvjpegStatus_t (*EncoderParamsSetOptimizedHuffman)(nvjpegEncoderParams_t, const int, cudaStream_t);
    nvjpegStatus_t (*EncoderParamsSetSamplingFactors)(nvjpegEncoderParams_t, const nvjpegChromaSubsampling_t, cudaStream_t);
    nvjpegStatus_t (*EncodeImage)(nvjpegHandle_t, nvjpegEncoderState_t, const nvjpegEncoderParams_t, const nvjpegImage_t *, nvjpegInputFormat_t, int, int,
                                  cudaStream_t);
    nvjpegStatus_t (*EncodeRetrieveBitstream)(nvjpegHandle_t, nvjpegEncoderState_t, unsigned char *, size_t *, cudaStream_t);
};

struct JpegState {
    JpegApi api;
    nvjpegHandle_t handle;
    nvjpegJpegState_t dec;
    nvjpegEncoderState_t enc;
    nvjpegEncoderParams_t params;
};

inline int jpeg_fail(const char *what, int status)
{
    char num[32];
    snprintf(num, sizeof num, "%d", status);
    return fail(B200GEO_ERR_CUDA, "nvjpeg: %s failed with status %s", what, num);
}

#define B2G_JPEG(st, call, ...)                                                 \
    do {                                                                        \
        nvjpegStatus_t s__ = (st)->api.call(__VA_ARGS__);                       \
        if (s__ != NVJPEG_STATUS_SUCCESS) return jpeg_fail(#call, (int)s__);    \
    } while (0)

// the context's codec state, created on first use (the caller holds the device)
inline int jpeg_state(b200geo_ctx *ctx, JpegState **out)
{
    if (ctx->jpeg) { *out = static_cast<JpegState *>(ctx->jpeg); return B200GEO_OK; }
    JpegState *st = new (std::nothrow) JpegState();
    if (!st) return fail(B200GEO_ERR_ARG, "out of host memory%s%s");
    memset(st, 0, sizeof *st);
    const char *names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
    for (const char *n : names)
        if ((st->api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!st->api.lib) { delete st; return fail(B200GEO_ERR_NODEVICE, "libnvjpeg.so.12 cannot be loaded: %s%s", dlerror()); }
#define B2G_SYM(field, name)                                                                                     \
    if (!(*reinterpret_cast<void **>(&st->api.field) = dlsym(st->api.lib, name))) {                              \
        delete st;                                                                                               \
        return fail(B200GEO_ERR_NODEVICE, "libnvjpeg lacks %s%s", name);                                         \
    }
    B2G_SYM(CreateSimple, "nvjpegCreateSimple") B2G_SYM(Destroy, "nvjpegDestroy")
    B2G_SYM(JpegStateCreate, "nvjpegJpegStateCreate") B2G_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy")
    B2G_SYM(GetImageInfo, "nvjpegGetImageInfo") B2G_SYM(Decode, "nvjpegDecode")
    B2G_SYM(EncoderStateCreate, "nvjpegEncoderStateCreate") B2G_SYM(EncoderStateDestroy, "nvjpegEncoderStateDestroy")
    B2G_SYM(EncoderParamsCreate, "nvjpegEncoderParamsCreate") B2G_SYM(EncoderParamsDestroy, "nvjpegEncoderParamsDestroy")
    B2G_SYM(EncoderParamsSetQuality, "nvjpegEncoderParamsSetQuality")
    B2G_SYM(EncoderParamsSetOptimizedHuffman, "nvjpegEncoderParamsSetOptimizedHuffman")
    B2G_SYM(EncoderParamsSetSamplingFactors, "nvjpegEncoderParamsSetSamplingFactors")
    B2G_SYM(EncodeImage, "nvjpegEncodeImage") B2G_SYM(EncodeRetrieveBitstream, "nvjpegEncodeRetrieveBitstream")
#undef B2G_SYM
    nvjpegStatus_t s;
    if ((s = st->api.CreateSimple(&st->handle)) != NVJPEG_STATUS_SUCCESS) { delete st; return jpeg_fail("CreateSimple", (int)s); }
    if ((s = st->api.JpegStateCreate(st->handle, &st->dec)) != NVJPEG_STATUS_SUCCESS) { delete st; return jpeg_fail("JpegStateCreate", (int)s); }
    if ((s = st->api.EncoderStateCreate(st->handle, &st->enc, 0)) != NVJPEG_STATUS_SUCCESS) { delete st; return jpeg_fail("EncoderStateCreate", (int)s); }
    if ((s = st->api.EncoderParamsCreate(st->handle, &st->params, 0)) != NVJPEG_STATUS_SUCCESS) { delete st; return jpeg_fail("EncoderParamsCreate", (int)s); }
    ctx->jpeg = st;
    *out = st;
    return B200GEO_OK;
}

inline void jpeg_release(b200geo_ctx *ctx)
{
    JpegState *st = static_cast<JpegState *>(ctx->jpeg);
    if (!st) return;
    if (st->params) st->api.EncoderParamsDestroy(st->params);
    if (st->enc) st->api.EncoderStateDestroy(st->enc);
    if (st->dec) st->api.JpegStateDestroy(st->dec);
    if (st->handle) st->api.Destroy(st->handle);
    delete st;                                 // (the library stays mapped: unloading a CUDA library at exit is not safe)
    ctx->jpeg = nullptr;
}

}  // namespace b200geo
