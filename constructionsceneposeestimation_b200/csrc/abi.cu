// Library-level entry points of libcspe.so: version, thread-local error string, device info.
#include <stdarg.h>
#include <string.h>

#include "cspe_common.cuh"

namespace cspe {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// SM count of the current device (immutable attribute; cached per device ordinal).
int sm_count() {
  static int cache[64] = {0};
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return -1;
  if (dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (dev < 64) cache[dev] = n;
  return n;
}

}  // namespace cspe

extern "C" int cspe_version(void) { return CSPE_ABI_VERSION; }

extern "C" const char* cspe_last_error(void) { return cspe::g_error; }

extern "C" int cspe_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0) {
    cspe::set_error("cspe_device_info: no current CUDA device (%s)", cudaGetErrorString(e));
    return CSPE_ERR_NO_DEVICE;
  }
  int n = 0, maj = 0, min = 0;
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  CSPE_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return CSPE_OK;
}

// Stream-ordered copy through the library's own runtime (cudaMemcpyDefault): lets a host driver put the D2H
// read-back of a batch into a captured CUDA graph next to the kernels without going through a framework's
// pinned-memory bookkeeping.  Host memory should be page-locked or the copy is not asynchronous.
extern "C" int cspe_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return CSPE_OK;
  CSPE_REQUIRE(dst != nullptr && src != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_memcpy_async: null pointer");
  CSPE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return CSPE_OK;
}

// Diagnostics for captured step graphs: how many dependency edges the graph holds and how many of them are
// PROGRAMMATIC (a kernel launched with programmatic stream serialisation behind another kernel keeps its
// early-start edge through stream capture; anything else between two kernels turns it into a full edge).
extern "C" int cspe_graph_edge_kinds(void* cuda_graph, int* num_nodes, int* num_edges, int* num_programmatic) {
  CSPE_REQUIRE(cuda_graph != nullptr, CSPE_ERR_INVALID_ARGUMENT, "cspe_graph_edge_kinds: graph is null");
  cudaGraph_t g = static_cast<cudaGraph_t>(cuda_graph);
  size_t nn = 0, ne = 0;
  CSPE_CUDA_OK(cudaGraphGetNodes(g, nullptr, &nn));
  CSPE_CUDA_OK(cudaGraphGetEdges_v2(g, nullptr, nullptr, nullptr, &ne));
  int prog = 0;
  if (ne > 0) {
    cudaGraphNode_t* from = new cudaGraphNode_t[ne];
    cudaGraphNode_t* to = new cudaGraphNode_t[ne];
    cudaGraphEdgeData* data = new cudaGraphEdgeData[ne];
    const cudaError_t e = cudaGraphGetEdges_v2(g, from, to, data, &ne);
    if (e == cudaSuccess)
      for (size_t i = 0; i < ne; ++i) prog += data[i].type == cudaGraphDependencyTypeProgrammatic ? 1 : 0;
    delete[] from;
    delete[] to;
    delete[] data;
    CSPE_CUDA_OK(e);
  }
  if (num_nodes) *num_nodes = static_cast<int>(nn);
  if (num_edges) *num_edges = static_cast<int>(ne);
  if (num_programmatic) *num_programmatic = prog;
  return CSPE_OK;
}
