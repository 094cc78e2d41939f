// api.cu -- context management and small utilities of the C ABI (include/ofb200.h).
#include "common.cuh"

static thread_local char g_err[1024] = "";

void ofb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ofb_last_error(void) { return g_err; }
extern "C" int ofb_version(void) { return 100; }

static int ctx_init(int device, cudaStream_t stream, bool own, ofb_ctx** out)
{
    OFB_REQUIRE(out, "ctx_create: null output");
    int ndev = 0;
    OFB_CUDA(cudaGetDeviceCount(&ndev));
    OFB_REQUIRE(device >= 0 && device < ndev, "ctx_create: device %d not available (%d visible)", device, ndev);
    OFB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ofb_set_error("ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return OFB_E_UNSUPPORTED;
    }
    ofb_ctx* c = new ofb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (own) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { ofb_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete c; return OFB_E_CUDA; }
        c->own_stream = true;
    } else {
        c->stream = stream;
        c->own_stream = false;
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    *out = c;
    return OFB_OK;
}

extern "C" int ofb_ctx_create(int device, ofb_ctx** out) { return ctx_init(device, nullptr, true, out); }

extern "C" int ofb_ctx_create_on_stream(int device, void* cuda_stream, ofb_ctx** out)
{
    return ctx_init(device, (cudaStream_t)cuda_stream, false, out);
}

extern "C" int ofb_ctx_destroy(ofb_ctx* ctx)
{
    OFB_REQUIRE(ctx, "ctx_destroy: null context");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->twin) { ofb_ctx_destroy(ctx->twin); cudaSetDevice(ctx->device); }
    if (ctx->ev_twin_fork) cudaEventDestroy(ctx->ev_twin_fork);
    if (ctx->ev_twin_join) cudaEventDestroy(ctx->ev_twin_join);
    for (ofb_pyr* p : ctx->pyramids) { cudaFree(p->base); delete p; }
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < 2; ++i) if (ctx->pair_pyr[s][i]) { cudaFree(ctx->pair_pyr[s][i]->base); delete ctx->pair_pyr[s][i]; }
    for (int s = 0; s < 2; ++s) {
        if (ctx->ev_ready[s]) cudaEventDestroy(ctx->ev_ready[s]);
        if (ctx->ev_free[s]) cudaEventDestroy(ctx->ev_free[s]);
    }
    for (cudaEvent_t e : ctx->ev_piece) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux2_stream) cudaStreamDestroy(ctx->aux2_stream);
    for (int i = 0; i < OFB_NSCRATCH; ++i) ctx->scratch[i].release();
    for (int i = 0; i < 4; ++i) ctx->pin[i].release();
    for (int i = 0; i < OFB_NSTAGE_EV; ++i) if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return OFB_OK;
}

extern "C" int ofb_ctx_sync(ofb_ctx* ctx)
{
    OFB_REQUIRE(ctx, "ctx_sync: null context");
    OFB_CUDA(ofb_join_aux(ctx));
    OFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return OFB_OK;
}

extern "C" int ofb_ctx_stream(ofb_ctx* ctx, void** cuda_stream_out)
{
    OFB_REQUIRE(ctx && cuda_stream_out, "ctx_stream: null argument");
    *cuda_stream_out = (void*)ctx->stream;
    ctx->stream_exported = true;
    return OFB_OK;
}

extern "C" int ofb_ctx_launch_count(ofb_ctx* ctx, uint64_t* out)
{
    OFB_REQUIRE(ctx && out, "ctx_launch_count: null argument");
    *out = ctx->launches;
    return OFB_OK;
}

extern "C" int ofb_timer_start(ofb_ctx* ctx)
{
    OFB_REQUIRE(ctx, "timer_start: null context");
    OFB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return OFB_OK;
}

extern "C" int ofb_timer_stop(ofb_ctx* ctx, float* ms_out)
{
    OFB_REQUIRE(ctx && ms_out, "timer_stop: null argument");
    OFB_CUDA(ofb_join_aux(ctx));
    OFB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    OFB_CUDA(cudaEventSynchronize(ctx->ev1));
    OFB_CUDA(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
    return OFB_OK;
}

extern "C" int ofb_dev_alloc(ofb_ctx* ctx, size_t bytes, void** out)
{
    OFB_REQUIRE(ctx && out && bytes > 0, "dev_alloc: bad argument");
    OFB_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) { ofb_set_error("dev_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return OFB_E_NOMEM; }
    return OFB_OK;
}
extern "C" int ofb_dev_free(ofb_ctx* ctx, void* p)
{
    OFB_REQUIRE(ctx, "dev_free: null context");
    if (p) { OFB_CUDA(cudaStreamSynchronize(ctx->stream)); OFB_CUDA(cudaFree(p)); }
    return OFB_OK;
}
extern "C" int ofb_host_alloc_pinned(ofb_ctx* ctx, size_t bytes, void** out)
{
    OFB_REQUIRE(ctx && out && bytes > 0, "host_alloc_pinned: bad argument");
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) { ofb_set_error("host_alloc_pinned(%zu): %s", bytes, cudaGetErrorString(e)); return OFB_E_NOMEM; }
    return OFB_OK;
}
extern "C" int ofb_host_free_pinned(ofb_ctx* ctx, void* p)
{
    OFB_REQUIRE(ctx, "host_free_pinned: null context");
    if (p) OFB_CUDA(cudaFreeHost(p));
    return OFB_OK;
}
extern "C" int ofb_memcpy_async(ofb_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    OFB_REQUIRE(ctx && dst && src, "memcpy: null argument");
    OFB_CUDA(ofb_join_aux(ctx));
    OFB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    if (ofb_is_device_ptr(dst)) ctx->async_writes++;
    return OFB_OK;
}
extern "C" int ofb_memcpy(ofb_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    OFB_TRY(ofb_memcpy_async(ctx, dst, src, bytes));
    OFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return OFB_OK;
}

extern "C" int ofb_ctx_set_profile(ofb_ctx* ctx, int enable)
{
    OFB_REQUIRE(ctx, "ctx_set_profile: null context");
    if (enable && !ctx->stage_ev[0])
        for (int i = 0; i < OFB_NSTAGE_EV; ++i) OFB_CUDA(cudaEventCreate(&ctx->stage_ev[i]));
    ctx->profile = enable != 0;
    for (int i = 0; i < OFB_NSTAGES; ++i) ctx->stage_ms[i] = 0.f;
    ctx->stage_calls = 0;
    return OFB_OK;
}

extern "C" int ofb_ctx_stage_times(ofb_ctx* ctx, float* ms_out, uint64_t* calls_out)
{
    OFB_REQUIRE(ctx && ms_out, "ctx_stage_times: null argument");
    for (int i = 0; i < OFB_NSTAGES; ++i) ms_out[i] = ctx->stage_ms[i];
    if (calls_out) *calls_out = ctx->stage_calls;
    return OFB_OK;
}
