// Context, device scratch arena, host<->device staging of API images.
#include <cstdarg>

#include "common.cuh"

int ds_fail(docscan_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

extern "C" int docscan_version(void) { return 100; }

extern "C" const char* docscan_strerror(int code) {
    switch (code) {
        case DOCSCAN_OK: return "ok";
        case DOCSCAN_ERR_NO_DEVICE: return "no CUDA device (libdocscan has no CPU fallback)";
        case DOCSCAN_ERR_CUDA: return "CUDA error";
        case DOCSCAN_ERR_BAD_ARG: return "bad argument";
        case DOCSCAN_ERR_NOMEM: return "out of memory";
        case DOCSCAN_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown error";
    }
}

extern "C" int docscan_create(int device, void* stream, docscan_ctx** out) {
    if (!out) return DOCSCAN_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return DOCSCAN_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return DOCSCAN_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DOCSCAN_ERR_CUDA;
    if (prop.major < 10) return DOCSCAN_ERR_UNSUPPORTED;   // sm_100a cubins only
    docscan_ctx* ctx = new docscan_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return DOCSCAN_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    ctx->pinned_size = 1 << 20;
    if (cudaHostAlloc((void**)&ctx->pinned, ctx->pinned_size, cudaHostAllocDefault) != cudaSuccess) {
        delete ctx;
        return DOCSCAN_ERR_NOMEM;
    }
    *out = ctx;
    return DOCSCAN_OK;
}

static void free_retired(docscan_ctx* ctx) {
    for (uint8_t* p : ctx->retired) cudaFree(p);
    ctx->retired.clear();
}

extern "C" int docscan_destroy(docscan_ctx* ctx) {
    if (!ctx) return DOCSCAN_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_retired(ctx);
    for (auto& kv : ctx->tables) cudaFree(kv.second);
    for (auto& kv : ctx->tc_tables) cudaFree(kv.second);
    for (void* p : ctx->user_allocs) cudaFree(p);
    if (ctx->angles_dev) cudaFree(ctx->angles_dev);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->tc_status) cudaFreeHost(ctx->tc_status);
    if (ctx->pin_mirror) cudaFreeHost(ctx->pin_mirror);
    if (ctx->copy_in) {
        cudaStreamDestroy(ctx->copy_in);
        cudaStreamDestroy(ctx->copy_out);
        for (cudaEvent_t e : ctx->pipe_ev) cudaEventDestroy(e);
    }
    for (int k = 0; k < DS_MAX_STREAMS; k++) {
        if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
        if (ctx->aux_ev[k]) cudaEventDestroy(ctx->aux_ev[k]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DOCSCAN_OK;
}

extern "C" int docscan_sync(docscan_ctx* ctx) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_retired(ctx);
    return DOCSCAN_OK;
}

extern "C" const char* docscan_last_error(docscan_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" int docscan_get_stream(docscan_ctx* ctx, void** stream) {
    if (!ctx || !stream) return DOCSCAN_ERR_BAD_ARG;
    *stream = (void*)ctx->stream;
    return DOCSCAN_OK;
}

extern "C" int64_t docscan_launch_count(docscan_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int docscan_transfer_bytes(docscan_ctx* ctx, int64_t* h2d, int64_t* d2h) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
    return DOCSCAN_OK;
}

extern "C" int docscan_profile_enable(docscan_ctx* ctx, int on) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    ctx->prof_on = on != 0;
    return DOCSCAN_OK;
}

extern "C" int docscan_profile_dump(docscan_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return DOCSCAN_ERR_BAD_ARG;
    cudaStreamSynchronize(ctx->stream);
    struct Agg { long n = 0; double ms = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        Agg& a = agg[r.name];
        a.n++; a.ms += ms; a.bytes += r.bytes;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    size_t off = 0;
    for (auto& kv : agg) {
        int w = snprintf(buf + off, cap - off, "%s %ld %.6f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes);
        if (w < 0 || (size_t)w >= cap - off) break;
        off += (size_t)w;
    }
    buf[off < cap ? off : cap - 1] = 0;
    return (int)off;
}

extern "C" int docscan_host_alloc(docscan_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return DOCSCAN_ERR_BAD_ARG;
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    DS_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return DOCSCAN_OK;
}
extern "C" int docscan_host_free(docscan_ctx* ctx, void* p) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    if (p) DS_CUDA(ctx, cudaFreeHost(p));
    return DOCSCAN_OK;
}
extern "C" int docscan_device_alloc(docscan_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return DOCSCAN_ERR_BAD_ARG;
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    DS_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    ctx->user_allocs.push_back(*out);
    return DOCSCAN_OK;
}
extern "C" int docscan_device_free(docscan_ctx* ctx, void* p) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    if (!p) return DOCSCAN_OK;
    for (size_t i = 0; i < ctx->user_allocs.size(); i++)
        if (ctx->user_allocs[i] == p) {
            ctx->user_allocs.erase(ctx->user_allocs.begin() + i);
            break;
        }
    DS_CUDA(ctx, cudaFree(p));
    return DOCSCAN_OK;
}
extern "C" int docscan_memcpy_h2d(docscan_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    DS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DOCSCAN_OK;
}
extern "C" int docscan_memcpy_d2h(docscan_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    DS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DOCSCAN_OK;
}

// ---- arena --------------------------------------------------------------------------------------
int ds_arena_reserve(docscan_ctx* ctx, size_t total) {
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    total += 1 << 16;
    if (ctx->arena_off + total <= ctx->arena_size) return DOCSCAN_OK;
    if (ctx->arena_off != 0)
        return ds_fail(ctx, DOCSCAN_ERR_NOMEM, "internal: arena reserve inside a live scope");
    // earlier calls may still be running on the old block: retire it, free at the next sync point
    if (ctx->arena) ctx->retired.push_back(ctx->arena);
    ctx->arena = nullptr;
    size_t want = total + total / 4;
    cudaError_t e = cudaMalloc((void**)&ctx->arena, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        free_retired(ctx);
        want = total;
        e = cudaMalloc((void**)&ctx->arena, want);
    }
    if (e != cudaSuccess) {
        ctx->arena = nullptr;
        ctx->arena_size = 0;
        return ds_fail(ctx, DOCSCAN_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    ctx->arena_size = want;
    return DOCSCAN_OK;
}

int ds_arena_alloc(docscan_ctx* ctx, size_t bytes, void** out) {
    size_t off = (ctx->arena_off + 255) & ~(size_t)255;
    if (off + bytes > ctx->arena_size)
        return ds_fail(ctx, DOCSCAN_ERR_NOMEM, "internal: device scratch arena too small (%zu + %zu > %zu)", off,
                       bytes, ctx->arena_size);
    *out = ctx->arena + off;
    ctx->arena_off = off + bytes;
    return DOCSCAN_OK;
}

int ds_arena_image(docscan_ctx* ctx, int w, int h, int ch, DImg* out) {
    size_t pitch = ((size_t)w * ch + 127) & ~(size_t)127;
    void* p = nullptr;
    DS_TRY(ds_arena_alloc(ctx, pitch * (size_t)h + 256, &p));
    out->p = (uint8_t*)p;
    out->w = w; out->h = h; out->pitch = (int)pitch; out->ch = ch;
    return DOCSCAN_OK;
}

int ds_pinned_alloc(docscan_ctx* ctx, size_t bytes, void** out) {
    size_t off = (ctx->pinned_off + 63) & ~(size_t)63;
    if (off + bytes > ctx->pinned_size) return ds_fail(ctx, DOCSCAN_ERR_NOMEM, "internal: pinned staging too small");
    *out = ctx->pinned + off;
    ctx->pinned_off = off + bytes;
    return DOCSCAN_OK;
}

int ds_upload(docscan_ctx* ctx, const void* host, size_t bytes, void** dev_out) {
    DS_TRY(ds_arena_alloc(ctx, bytes, dev_out));
    // pageable source: the runtime snapshots it before returning, so `host` may be reused at once
    DS_CUDA(ctx, cudaMemcpyAsync(*dev_out, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return DOCSCAN_OK;
}

// ---- API image staging ----------------------------------------------------------------------------
int ds_check_image(docscan_ctx* ctx, const docscan_image* im, int channels, const char* what) {
    if (!im || !im->data) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: null image", what);
    if (im->width <= 0 || im->height <= 0) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: empty image", what);
    if (channels && im->channels != channels)
        return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: expected %d channel(s), got %d", what, channels, im->channels);
    if (im->channels != 1 && im->channels != 3)
        return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: channels must be 1 or 3", what);
    if ((int64_t)im->pitch < (int64_t)im->width * im->channels)
        return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: pitch smaller than a row", what);
    if (im->space != DOCSCAN_HOST && im->space != DOCSCAN_DEVICE)
        return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: bad memory space", what);
    return DOCSCAN_OK;
}

int ds_stage_in(docscan_ctx* ctx, const docscan_image* im, DImg* out) {
    if (im->space == DOCSCAN_DEVICE) {
        out->p = (uint8_t*)im->data; out->w = im->width; out->h = im->height; out->pitch = im->pitch; out->ch = im->channels;
        return DOCSCAN_OK;
    }
    DS_TRY(ds_arena_image(ctx, im->width, im->height, im->channels, out));
    ctx->h2d_bytes += (int64_t)im->width * im->channels * im->height;
    DS_CUDA(ctx, cudaMemcpy2DAsync(out->p, out->pitch, im->data, im->pitch, (size_t)im->width * im->channels,
                                   im->height, cudaMemcpyHostToDevice, ctx->stream));
    return DOCSCAN_OK;
}

int ds_stage_out_begin(docscan_ctx* ctx, const docscan_image* im, DImg* out) {
    if (im->space == DOCSCAN_DEVICE) {
        out->p = (uint8_t*)im->data; out->w = im->width; out->h = im->height; out->pitch = im->pitch; out->ch = im->channels;
        return DOCSCAN_OK;
    }
    return ds_arena_image(ctx, im->width, im->height, im->channels, out);
}

int ds_stage_out_end(docscan_ctx* ctx, const docscan_image* im, const DImg& dev) {
    if (im->space == DOCSCAN_DEVICE) return DOCSCAN_OK;
    ctx->d2h_bytes += (int64_t)im->width * im->channels * im->height;
    DS_CUDA(ctx, cudaMemcpy2DAsync(im->data, im->pitch, dev.p, dev.pitch, (size_t)im->width * im->channels,
                                   im->height, cudaMemcpyDeviceToHost, ctx->stream));
    return DOCSCAN_OK;
}

int ds_finish(docscan_ctx* ctx, bool any_host) {
    if (any_host) {
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        free_retired(ctx);
    }
    return DOCSCAN_OK;
}
