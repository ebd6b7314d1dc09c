// Weight packing (state_dict layout <-> kernel layout), patch extraction / reassembly index
// kernels, and the library's bookkeeping entry points.
#include <stdarg.h>
#include <string.h>
#include "dpi_common.cuh"

namespace dpi {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// w [Cout_l][Cin_l][taps]  ->  w_fwd [Cout_p][taps][Cin_p],  w_dgrad [Cin_p][taps][Cout_p]
__global__ void pack_weights_kernel(const float* __restrict__ w, const int32_t* __restrict__ cout_map,
                                    const int32_t* __restrict__ cin_map, int Cout_l, int Cin_l, int Cout_p,
                                    int Cin_p, int taps, float* __restrict__ w_fwd, float* __restrict__ w_dgrad,
                                    const float* __restrict__ bias, float* __restrict__ bias_packed, int rtf32) {
  if (bias_packed && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < Cout_p; c += blockDim.x) {
      const int lo = cout_map ? cout_map[c] : c;
      bias_packed[c] = (bias && lo >= 0 && lo < Cout_l) ? bias[lo] : 0.f;
    }
  }
  const int64_t total = (int64_t)Cout_p * taps * Cin_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int t = (int)((i / Cin_p) % taps);
    const int co = (int)(i / ((int64_t)Cin_p * taps));
    const int lo = cout_map ? cout_map[co] : co, li = cin_map ? cin_map[ci] : ci;
    float v = 0.f;
    if (lo >= 0 && lo < Cout_l && li >= 0 && li < Cin_l) v = w[((int64_t)lo * Cin_l + li) * taps + t];
    if (rtf32) v = round_tf32(v);
    if (w_fwd) w_fwd[i] = v;
    if (w_dgrad) w_dgrad[((int64_t)ci * taps + t) * Cout_p + co] = v;
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, const int32_t* __restrict__ cout_map,
                                    const int32_t* __restrict__ cin_map, int Cout_l, int Cin_l, int Cout_p,
                                    int Cin_p, int taps, float* __restrict__ dw) {
  const int64_t total = (int64_t)Cout_p * taps * Cin_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int t = (int)((i / Cin_p) % taps);
    const int co = (int)(i / ((int64_t)Cin_p * taps));
    const int lo = cout_map ? cout_map[co] : co, li = cin_map ? cin_map[ci] : ci;
    if (lo >= 0 && lo < Cout_l && li >= 0 && li < Cin_l) dw[((int64_t)lo * Cin_l + li) * taps + t] = dwp[i];
  }
}

// ---- batched forms: one launch for all conv layers of a network (grid.y = layer) --------------------
__global__ void pack_weights_batched_kernel(const dpi_pack_job* __restrict__ jobs, int rtf32) {
  const dpi_pack_job j = jobs[blockIdx.y];
  if (j.bias_packed && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < j.Cout_p; c += blockDim.x) {
      const int lo = j.cout_map ? j.cout_map[c] : c;
      j.bias_packed[c] = (j.bias && lo >= 0 && lo < j.Cout_l) ? j.bias[lo] : 0.f;
    }
  }
  const int total = j.Cout_p * j.taps * j.Cin_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % j.Cin_p;
    const int t = (i / j.Cin_p) % j.taps;
    const int co = i / (j.Cin_p * j.taps);
    const int lo = j.cout_map ? j.cout_map[co] : co, li = j.cin_map ? j.cin_map[ci] : ci;
    float v = 0.f;
    if (lo >= 0 && lo < j.Cout_l && li >= 0 && li < j.Cin_l) v = j.w[((int64_t)lo * j.Cin_l + li) * j.taps + t];
    if (rtf32) v = round_tf32(v);
    if (j.w_fwd) j.w_fwd[i] = v;
    if (j.w_dgrad) j.w_dgrad[((int64_t)ci * j.taps + t) * j.Cout_p + co] = v;
    if (j.w_dgrad_cat) j.w_dgrad_cat[((int64_t)ci * j.cat_taps + j.cat_tap0 + t) * j.cat_ld + j.cat_off + co] = v;
  }
}

__global__ void unpack_wgrad_batched_kernel(const dpi_pack_job* __restrict__ jobs) {
  const dpi_pack_job j = jobs[blockIdx.y];
  const int total = j.Cout_p * j.taps * j.Cin_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % j.Cin_p;
    const int t = (i / j.Cin_p) % j.taps;
    const int co = i / (j.Cin_p * j.taps);
    const int lo = j.cout_map ? j.cout_map[co] : co, li = j.cin_map ? j.cin_map[ci] : ci;
    if (lo >= 0 && lo < j.Cout_l && li >= 0 && li < j.Cin_l) j.dw[((int64_t)lo * j.Cin_l + li) * j.taps + t] = j.dw_packed[i];
  }
}

// ---- patches -------------------------------------------------------------------------------------
struct PatchGeom {
  int vs[3], ps[3], st[3], np[3];
};
static int make_patch_geom(const int32_t* vol_shape3, const int32_t* patch_shape3, const int32_t* stride3,
                           PatchGeom& g) {
  for (int a = 0; a < 3; ++a) {
    g.vs[a] = vol_shape3[a]; g.ps[a] = patch_shape3[a]; g.st[a] = stride3[a];
    if (g.vs[a] < 1 || g.ps[a] < 1 || g.st[a] < 1 || g.ps[a] > g.vs[a]) {
      set_error("patch geometry: axis %d volume %d patch %d stride %d", a, g.vs[a], g.ps[a], g.st[a]);
      return DPI_ERR_INVALID_ARG;
    }
    g.np[a] = (g.vs[a] - g.ps[a]) / g.st[a] + 1;
  }
  return DPI_OK;
}

__global__ void patch_extract_kernel(const double* __restrict__ vol, PatchGeom g, double gain,
                                     double* __restrict__ patches) {
  const int64_t pvox = (int64_t)g.ps[0] * g.ps[1] * g.ps[2];
  const int64_t total = pvox * g.np[0] * g.np[1] * g.np[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int z = (int)(r % g.ps[2]); r /= g.ps[2];
    const int y = (int)(r % g.ps[1]); r /= g.ps[1];
    const int x = (int)(r % g.ps[0]); r /= g.ps[0];
    const int pz = (int)(r % g.np[2]); r /= g.np[2];
    const int py = (int)(r % g.np[1]);
    const int px = (int)(r / g.np[1]);
    const int64_t src = ((int64_t)(px * g.st[0] + x) * g.vs[1] + (py * g.st[1] + y)) * g.vs[2] + (pz * g.st[2] + z);
    patches[i] = vol[src] * gain;
  }
}

// gather form of the overlap-add: every output voxel sums the patches that cover it in increasing
// patch index (the order of the nested NumPy loops), in fp64, then /count -> float32 -> /gain.
__global__ void patch_reassemble_kernel(const float* __restrict__ patches, PatchGeom g, float gain,
                                        float* __restrict__ vol) {
  int cs[3];
  for (int a = 0; a < 3; ++a) cs[a] = (g.np[a] - 1) * g.st[a] + g.ps[a];
  const int64_t pvox = (int64_t)g.ps[0] * g.ps[1] * g.ps[2];
  const int64_t total = (int64_t)cs[0] * cs[1] * cs[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c2 = (int)(r % cs[2]); r /= cs[2];
    const int c1 = (int)(r % cs[1]);
    const int c0 = (int)(r / cs[1]);
    const int co[3] = {c0, c1, c2};
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      // patches p with p*st <= c < p*st + ps
      int l = co[a] - g.ps[a] + 1;
      l = l <= 0 ? 0 : (l + g.st[a] - 1) / g.st[a];
      int h = co[a] / g.st[a];
      if (h > g.np[a] - 1) h = g.np[a] - 1;
      lo[a] = l; hi[a] = h;
    }
    double acc = 0.0, cnt = 0.0;
    for (int p0 = lo[0]; p0 <= hi[0]; ++p0)
      for (int p1 = lo[1]; p1 <= hi[1]; ++p1)
        for (int p2 = lo[2]; p2 <= hi[2]; ++p2) {
          const int64_t pidx = ((int64_t)p0 * g.np[1] + p1) * g.np[2] + p2;
          const int64_t off = ((int64_t)(c0 - p0 * g.st[0]) * g.ps[1] + (c1 - p1 * g.st[1])) * g.ps[2] + (c2 - p2 * g.st[2]);
          acc += (double)patches[pidx * pvox + off];
          cnt += 1.0;
        }
    const float avg = (float)(acc / cnt);
    vol[i] = avg / gain;
  }
}

}  // namespace dpi

using namespace dpi;

extern "C" {

const char* dpi_last_error_string(void) { return g_err; }
int dpi_version(void) { return 100; }
int64_t dpi_launch_count(void) { return g_launches.load(); }
int dpi_device_supports_tcgen05(int dev) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

int dpi_pack_conv_weights(const float* w, const int32_t* cout_map, const int32_t* cin_map, int Cout_l, int Cin_l,
                          int Cout_p, int Cin_p, int taps, float* w_fwd, float* w_dgrad, const float* bias,
                          float* bias_packed, int round_tf32, void* stream) {
  DPI_REQUIRE(w && (w_fwd || w_dgrad) && Cout_p > 0 && Cin_p > 0 && taps > 0, "dpi_pack_conv_weights: bad arguments");
  const int64_t total = (int64_t)Cout_p * taps * Cin_p;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, cout_map, cin_map, Cout_l, Cin_l, Cout_p, Cin_p,
                                                               taps, w_fwd, w_dgrad, bias, bias_packed, round_tf32);
  return check_launch("dpi_pack_conv_weights");
}

int dpi_unpack_conv_wgrad(const float* dw_packed, const int32_t* cout_map, const int32_t* cin_map, int Cout_l,
                          int Cin_l, int Cout_p, int Cin_p, int taps, float* dw, void* stream) {
  DPI_REQUIRE(dw_packed && dw && Cout_p > 0 && Cin_p > 0 && taps > 0, "dpi_unpack_conv_wgrad: bad arguments");
  const int64_t total = (int64_t)Cout_p * taps * Cin_p;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  unpack_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dw_packed, cout_map, cin_map, Cout_l, Cin_l, Cout_p,
                                                               Cin_p, taps, dw);
  return check_launch("dpi_unpack_conv_wgrad");
}

int dpi_pack_conv_weights_batched(const dpi_pack_job* jobs_dev, int njobs, int round_tf32, void* stream) {
  DPI_REQUIRE(jobs_dev && njobs > 0 && njobs <= 65535, "dpi_pack_conv_weights_batched: bad arguments");
  pack_weights_batched_kernel<<<dim3(48, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(jobs_dev, round_tf32);
  return check_launch("dpi_pack_conv_weights_batched");
}

int dpi_unpack_conv_wgrad_batched(const dpi_pack_job* jobs_dev, int njobs, void* stream) {
  DPI_REQUIRE(jobs_dev && njobs > 0 && njobs <= 65535, "dpi_unpack_conv_wgrad_batched: bad arguments");
  unpack_wgrad_batched_kernel<<<dim3(48, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(jobs_dev);
  return check_launch("dpi_unpack_conv_wgrad_batched");
}

int dpi_patch_extract_f64(const double* vol, const int32_t* vol_shape3, const int32_t* patch_shape3,
                          const int32_t* stride3, double gain, double* patches, void* stream) {
  DPI_REQUIRE(vol && patches && vol_shape3 && patch_shape3 && stride3, "dpi_patch_extract_f64: null pointer");
  PatchGeom g;
  int rc = make_patch_geom(vol_shape3, patch_shape3, stride3, g);
  if (rc) return rc;
  const int64_t total = (int64_t)g.ps[0] * g.ps[1] * g.ps[2] * g.np[0] * g.np[1] * g.np[2];
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  patch_extract_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(vol, g, gain, patches);
  return check_launch("dpi_patch_extract_f64");
}

int dpi_patch_reassemble_f32(const float* patches, const int32_t* vol_shape3, const int32_t* patch_shape3,
                             const int32_t* stride3, float gain, float* vol, void* stream) {
  DPI_REQUIRE(vol && patches && vol_shape3 && patch_shape3 && stride3, "dpi_patch_reassemble_f32: null pointer");
  PatchGeom g;
  int rc = make_patch_geom(vol_shape3, patch_shape3, stride3, g);
  if (rc) return rc;
  int64_t total = 1;
  for (int a = 0; a < 3; ++a) total *= (g.np[a] - 1) * g.st[a] + g.ps[a];
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  patch_reassemble_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(patches, g, gain, vol);
  return check_launch("dpi_patch_reassemble_f32");
}

}  // extern "C"
