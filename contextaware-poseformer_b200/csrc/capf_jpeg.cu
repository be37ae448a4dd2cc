// Frame decode for the dataset row (mvn/datasets/human36m.py:565-567: cv2.imread(path, IMREAD_COLOR) -> BGR uint8).
//
// JPEG decode is a LIBRARY call on both sides: libjpeg through OpenCV in the reference, nvJPEG (CUDA toolkit) here.  The
// point of doing it on the GPU is the data path: the compressed bytes are what crosses PCIe (~10x fewer than raw frames)
// and the decoded frame lands directly in the padded [B,Hs,Ws,3] storage CAPF_OP_WARP_AFFINE_U8 crops from, interleaved
// BGR like cv2's.  libnvjpeg is loaded lazily with dlopen, so libcapf_b200.so itself has no link-time dependency on it:
// on a box without nvJPEG every entry below returns CAPF_ERR_UNSUPPORTED and nothing else is affected.
// The two decoders are different implementations of the same standard (IDCT, chroma upsampling): pixels agree to a few
// grey levels, not bit for bit -- tests/test_dataset.py states the bound.
//
// Backends.  nvjpegDecode on the default backend decodes one frame per call with the Huffman stage on the calling host thread
// (measured: 411 frames/s for 1000x1000 frames, tools/jpeg_bench.py).  CAPF_JPEG_BACKEND=hardware | gpu_hybrid (read once, at the
// first decode) sends whole batches through nvjpegDecodeBatched on a second handle created with nvjpegCreateEx for that backend --
// the NVJPG engines, or Huffman decoding on the SMs.  Opt-in: these decoders round differently again, and a backend the
// GPU / driver / stream does not support is an error (no silent fallback), so that a measurement names what ran.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "capf_internal.h"

namespace {

struct NvJpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
  nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
  nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*, cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  nvjpegHandle_t bhandle = nullptr;          // batched decode on the backend CAPF_JPEG_BACKEND names
  nvjpegJpegState_t bstate = nullptr;
  int backend = -1;                          // -1: not read yet, 0: per-frame nvjpegDecode (default), else nvjpegBackend_t
  int device = -1;
  bool tried = false, ok = false;
};

NvJpegApi g_api;
std::mutex g_mu;

// caller holds g_mu
bool load_api() {
  if (g_api.tried) return g_api.ok;
  g_api.tried = true;
  const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
  for (const char* n : names) {
    g_api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib) return false;
  g_api.CreateSimple = reinterpret_cast<decltype(g_api.CreateSimple)>(dlsym(g_api.lib, "nvjpegCreateSimple"));
  g_api.JpegStateCreate = reinterpret_cast<decltype(g_api.JpegStateCreate)>(dlsym(g_api.lib, "nvjpegJpegStateCreate"));
  g_api.GetImageInfo = reinterpret_cast<decltype(g_api.GetImageInfo)>(dlsym(g_api.lib, "nvjpegGetImageInfo"));
  g_api.Decode = reinterpret_cast<decltype(g_api.Decode)>(dlsym(g_api.lib, "nvjpegDecode"));
  g_api.CreateEx = reinterpret_cast<decltype(g_api.CreateEx)>(dlsym(g_api.lib, "nvjpegCreateEx"));
  g_api.DecodeBatchedInitialize = reinterpret_cast<decltype(g_api.DecodeBatchedInitialize)>(dlsym(g_api.lib, "nvjpegDecodeBatchedInitialize"));
  g_api.DecodeBatched = reinterpret_cast<decltype(g_api.DecodeBatched)>(dlsym(g_api.lib, "nvjpegDecodeBatched"));
  g_api.ok = g_api.CreateSimple && g_api.JpegStateCreate && g_api.GetImageInfo && g_api.Decode;
  return g_api.ok;
}

// caller holds g_mu; the decoder state is bound to one device (one process per GPU)
int ensure_handle(int device) {
  if (!load_api()) return capf::set_error(CAPF_ERR_UNSUPPORTED, "jpeg: libnvjpeg could not be loaded");
  if (g_api.handle && g_api.device != device) return capf::set_error(CAPF_ERR_UNSUPPORTED, "jpeg: the decoder is bound to the device of its first use");
  if (g_api.handle) return CAPF_OK;
  if (cudaSetDevice(device) != cudaSuccess) return capf::set_error(CAPF_ERR_CUDA, "jpeg: cudaSetDevice failed");
  nvjpegStatus_t s = g_api.CreateSimple(&g_api.handle);
  if (s != NVJPEG_STATUS_SUCCESS) { g_api.handle = nullptr; return capf::set_errorf(CAPF_ERR_CUDA, "jpeg: nvjpegCreateSimple failed (%d)", (int)s); }
  s = g_api.JpegStateCreate(g_api.handle, &g_api.state);
  if (s != NVJPEG_STATUS_SUCCESS) { g_api.handle = nullptr; return capf::set_errorf(CAPF_ERR_CUDA, "jpeg: nvjpegJpegStateCreate failed (%d)", (int)s); }
  g_api.device = device;
  return CAPF_OK;
}

// caller holds g_mu, ensure_handle() succeeded.  Reads CAPF_JPEG_BACKEND once and creates the batched decoder it names.
int ensure_batched() {
  if (g_api.backend >= 0) return CAPF_OK;
  const char* e = std::getenv("CAPF_JPEG_BACKEND");
  int want = 0;
  if (e && *e && std::strcmp(e, "default") != 0) {
    if (!std::strcmp(e, "hardware")) want = (int)NVJPEG_BACKEND_HARDWARE;
    else if (!std::strcmp(e, "gpu_hybrid")) want = (int)NVJPEG_BACKEND_GPU_HYBRID;
    else if (!std::strcmp(e, "hybrid")) want = (int)NVJPEG_BACKEND_HYBRID;
    else return capf::set_errorf(CAPF_ERR_ARG, "jpeg: CAPF_JPEG_BACKEND=%s (default | hybrid | gpu_hybrid | hardware)", e);
  }
  if (want) {
    if (!g_api.CreateEx || !g_api.DecodeBatchedInitialize || !g_api.DecodeBatched)
      return capf::set_error(CAPF_ERR_UNSUPPORTED, "jpeg: this libnvjpeg has no batched decode entry points");
    nvjpegStatus_t s = g_api.CreateEx((nvjpegBackend_t)want, nullptr, nullptr, 0, &g_api.bhandle);
    if (s != NVJPEG_STATUS_SUCCESS) {
      g_api.bhandle = nullptr;
      return capf::set_errorf(CAPF_ERR_UNSUPPORTED, "jpeg: nvjpegCreateEx(backend %s) failed (status %d): not available on this GPU / driver", e, (int)s);
    }
    s = g_api.JpegStateCreate(g_api.bhandle, &g_api.bstate);
    if (s != NVJPEG_STATUS_SUCCESS) {
      g_api.bhandle = nullptr;
      return capf::set_errorf(CAPF_ERR_CUDA, "jpeg: nvjpegJpegStateCreate(backend %s) failed (%d)", e, (int)s);
    }
  }
  g_api.backend = want;
  return CAPF_OK;
}

}  // namespace

extern "C" {

int capf_jpeg_available(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return load_api() ? 1 : 0;
}

int capf_jpeg_info(const unsigned char* data, size_t length, int device, int* height, int* width) {
  if (!data || !length || !height || !width) return capf::set_error(CAPF_ERR_ARG, "jpeg_info: null argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (int rc = ensure_handle(device)) return rc;
  int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
  nvjpegChromaSubsampling_t css;
  const nvjpegStatus_t s = g_api.GetImageInfo(g_api.handle, data, length, &ncomp, &css, ws, hs);
  if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_ARG, "jpeg_info: not a decodable JPEG stream (nvjpeg status %d)", (int)s);
  *height = hs[0];
  *width = ws[0];
  return CAPF_OK;
}

int capf_jpeg_decode_batch(const unsigned char* const* data, const size_t* lengths, int n, unsigned char* frames, int Hs, int Ws, int* sizes_hw, int device,
                           void* stream) {
  if (!data || !lengths || n <= 0 || !frames || Hs <= 0 || Ws <= 0) return capf::set_error(CAPF_ERR_ARG, "jpeg_decode_batch: bad arguments");
  std::lock_guard<std::mutex> lk(g_mu);
  if (int rc = ensure_handle(device)) return rc;
  if (int rc = ensure_batched()) return rc;
  if (g_api.backend != 0) {
    std::vector<nvjpegImage_t> dst((size_t)n);
    for (int k = 0; k < n; ++k) {
      int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
      nvjpegChromaSubsampling_t css;
      const nvjpegStatus_t s = g_api.GetImageInfo(g_api.handle, data[k], lengths[k], &ncomp, &css, ws, hs);
      if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_ARG, "jpeg_decode_batch: stream %d is not a decodable JPEG (nvjpeg status %d)", k, (int)s);
      if (hs[0] > Hs || ws[0] > Ws) return capf::set_errorf(CAPF_ERR_ARG, "jpeg_decode_batch: frame %d is %dx%d, storage %dx%d", k, hs[0], ws[0], Hs, Ws);
      for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { dst[k].channel[c] = nullptr; dst[k].pitch[c] = 0; }
      dst[k].channel[0] = frames + (size_t)k * Hs * Ws * 3;
      dst[k].pitch[0] = (size_t)Ws * 3;
      if (sizes_hw) { sizes_hw[2 * k] = hs[0]; sizes_hw[2 * k + 1] = ws[0]; }
    }
    const unsigned hw = std::thread::hardware_concurrency();
    nvjpegStatus_t s = g_api.DecodeBatchedInitialize(g_api.bhandle, g_api.bstate, n, (int)(hw ? (hw < 16u ? hw : 16u) : 1u), NVJPEG_OUTPUT_BGRI);
    if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_CUDA, "jpeg_decode_batch: nvjpegDecodeBatchedInitialize failed (status %d)", (int)s);
    s = g_api.DecodeBatched(g_api.bhandle, g_api.bstate, data, lengths, dst.data(), (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_CUDA, "jpeg_decode_batch: nvjpegDecodeBatched failed (status %d)", (int)s);
    return CAPF_OK;
  }
  for (int k = 0; k < n; ++k) {
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
    nvjpegChromaSubsampling_t css;
    nvjpegStatus_t s = g_api.GetImageInfo(g_api.handle, data[k], lengths[k], &ncomp, &css, ws, hs);
    if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_ARG, "jpeg_decode_batch: stream %d is not a decodable JPEG (nvjpeg status %d)", k, (int)s);
    if (hs[0] > Hs || ws[0] > Ws) return capf::set_errorf(CAPF_ERR_ARG, "jpeg_decode_batch: frame %d is %dx%d, storage %dx%d", k, hs[0], ws[0], Hs, Ws);
    nvjpegImage_t dst;
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { dst.channel[c] = nullptr; dst.pitch[c] = 0; }
    dst.channel[0] = frames + (size_t)k * Hs * Ws * 3;
    dst.pitch[0] = (size_t)Ws * 3;
    s = g_api.Decode(g_api.handle, g_api.state, data[k], lengths[k], NVJPEG_OUTPUT_BGRI, &dst, (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) return capf::set_errorf(CAPF_ERR_CUDA, "jpeg_decode_batch: nvjpegDecode failed on stream %d (status %d)", k, (int)s);
    if (sizes_hw) { sizes_hw[2 * k] = hs[0]; sizes_hw[2 * k + 1] = ws[0]; }
  }
  return CAPF_OK;
}

}  // extern "C"
