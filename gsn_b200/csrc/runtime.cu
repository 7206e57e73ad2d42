// Host-side submission of one captured step (no kernels here).
//
// A step of the serving path is: one packed input copy (pinned host or device -> the bucket's static input buffer), one
// replay of the bucket's captured CUDA graph, optionally one copy of the predictions to pinned host memory -- all on the
// caller's stream.  Issued from Python as three framework calls (stream context + Tensor.copy_ + CUDAGraph.replay) a
// step costs 22-27 us of host time, which at B = 128 is the whole device time of the step (27 us with 8 steps in
// flight): the host was the limiter of the throughput measurement.  One C call does the same three driver calls.
#include "common.cuh"

using namespace gsn;

extern "C" int gsn_submit_step(void *graph_exec, void *d_in, const void *src, size_t in_bytes, void *h_out, const void *d_out,
                               size_t out_bytes, void *stream_) {
    if (!graph_exec || (in_bytes && (!d_in || !src)) || (out_bytes && (!h_out || !d_out))) return GSN_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (in_bytes) GSN_CUDA_OK(cudaMemcpyAsync(d_in, src, in_bytes, cudaMemcpyDefault, stream));
    GSN_CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, stream));
    if (out_bytes) GSN_CUDA_OK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
    return GSN_OK;
}
