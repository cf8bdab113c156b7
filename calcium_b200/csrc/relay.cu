// relay.cu -- host-side plumbing for the NVLink relay of the "final host gather" (SURVEY.md section 8e).
//
// On an 8-GPU box the device->host path is a shared platform resource and far from symmetric: with all
// eight ranks copying their encoded files out, the GPUs behind the slow root complex get ~11 GB/s each
// (profiles/r1_d2h_probe_8gpu.txt).  Routing their bytes over NVLink to a partner GPU whose host path is
// fast, and from there to pinned host memory, raises the aggregate from 93 to 176 GB/s
// (profiles/r2_relay_probe_8gpu.txt).  Ranks are separate processes (one per GPU), so the partner's
// staging ring is reached through CUDA IPC: the receiver allocates it with cudaMalloc (IPC handles cannot
// describe a caching allocator's sub-blocks), the sender maps it and pushes chunks with the copy engine
// (cudaMemcpyAsync peer copy -- no SM is involved, K1 owns all of them), ordering between the two
// processes' streams comes from interprocess events.  No kernels here; calcium_b200/relay.py drives it.
#include "common.cuh"

using namespace cab200;

extern "C" int cab200_relay_alloc(size_t bytes, void **dev_ptr_out, unsigned char handle_out[64])
{
    if (!bytes || !dev_ptr_out || !handle_out) { set_error("cab200_relay_alloc: bad argument"); return CAB200_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    CAB_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return CAB200_ECUDA;
    }
    memcpy(handle_out, &h, 64);
    *dev_ptr_out = p;
    return CAB200_OK;
}

extern "C" int cab200_relay_free(void *dev_ptr)
{
    if (dev_ptr) CAB_CUDA(cudaFree(dev_ptr));
    return CAB200_OK;
}

extern "C" int cab200_relay_open(const unsigned char handle[64], void **dev_ptr_out)
{
    if (!handle || !dev_ptr_out) { set_error("cab200_relay_open: bad argument"); return CAB200_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CAB_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return CAB200_OK;
}

extern "C" int cab200_relay_close(void *dev_ptr)
{
    if (dev_ptr) CAB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return CAB200_OK;
}

extern "C" int cab200_ipc_event_create(void **event_out, unsigned char handle_out[64])
{
    if (!event_out || !handle_out) { set_error("cab200_ipc_event_create: bad argument"); return CAB200_EINVAL; }
    static_assert(sizeof(cudaIpcEventHandle_t) == 64, "IPC handle size");
    cudaEvent_t ev;
    CAB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
    cudaIpcEventHandle_t h;
    cudaError_t e = cudaIpcGetEventHandle(&h, ev);
    if (e != cudaSuccess) {
        cudaEventDestroy(ev);
        set_error("cudaIpcGetEventHandle failed: %s", cudaGetErrorString(e));
        return CAB200_ECUDA;
    }
    memcpy(handle_out, &h, 64);
    *event_out = (void *)ev;
    return CAB200_OK;
}

extern "C" int cab200_ipc_event_open(const unsigned char handle[64], void **event_out)
{
    if (!handle || !event_out) { set_error("cab200_ipc_event_open: bad argument"); return CAB200_EINVAL; }
    cudaIpcEventHandle_t h;
    memcpy(&h, handle, 64);
    cudaEvent_t ev;
    CAB_CUDA(cudaIpcOpenEventHandle(&ev, h));
    *event_out = (void *)ev;
    return CAB200_OK;
}

extern "C" int cab200_event_create(void **event_out)
{
    if (!event_out) { set_error("cab200_event_create: bad argument"); return CAB200_EINVAL; }
    cudaEvent_t ev;
    CAB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    *event_out = (void *)ev;
    return CAB200_OK;
}

extern "C" int cab200_event_record(void *event, void *stream)
{
    CAB_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return CAB200_OK;
}

extern "C" int cab200_event_synchronize(void *event)
{
    CAB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return CAB200_OK;
}

extern "C" int cab200_event_destroy(void *event)
{
    if (event) CAB_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return CAB200_OK;
}

extern "C" int cab200_stream_wait_event(void *stream, void *event)
{
    CAB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return CAB200_OK;
}

extern "C" int cab200_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (!bytes) return CAB200_OK;
    if (!dst || !src) { set_error("cab200_copy_async: null pointer"); return CAB200_EINVAL; }
    CAB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return CAB200_OK;
}
