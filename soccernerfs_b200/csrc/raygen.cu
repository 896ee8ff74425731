// Pixel -> ray generation on the device (sm_100a): the perspective, undistorted case of
// Cameras._generate_rays_from_coords (NS/cameras/cameras.py:505-741) as used by RayGenerator.forward
// (NS/model_components/ray_generators.py:43-59, training: explicit (camera,row,col) triplets) and by
// Cameras.generate_rays(camera_indices=i, keep_shape=True) (cameras.py:327-502, evaluation: a row-major tile of a
// full frame).  One thread per ray; the reference's operation order is kept (no FMA contraction): image-plane
// coordinates ((x - cx) / fx, -(y - cy) / fy) of the pixel centre and of its +1 neighbours in x and y, rotation by
// the camera-to-world matrix as three products summed left to right, normalisation by max(norm, 4*DBL_EPSILON),
// pixel_area = |d - d_x| * |d - d_y|.  Fisheye / equirectangular cameras and non-zero distortion are rejected by the
// host wrapper (they raise), not approximated.
#include "common.cuh"

namespace kp {

struct RayGenArgs {
  const float* c2w;         // [n_cams,3,4]
  const float* intrinsics;  // [n_cams,4] = fx, fy, cx, cy
  const float* cam_times;   // [n_cams] or null
  const int64_t* ray_indices;  // [N,3] = camera,row,col or null (tile mode)
  int n_cams, cam, width;
  int64_t first_pixel, N;
  float pixel_offset;
  float* origins; float* directions; float* pixel_area; float* directions_norm; float* times;
};

__device__ __forceinline__ void rotate_normalize(const float* R, float dx, float dy, float dz, float out[3], float* norm_out) {
  float v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    v[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[i * 4 + 0]), __fmul_rn(dy, R[i * 4 + 1])), __fmul_rn(dz, R[i * 4 + 2]));
  float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
  n = fmaxf(n, 8.8817841970012523e-16f);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = __fdiv_rn(v[i], n);
  if (norm_out != nullptr) *norm_out = n;
}

__device__ __forceinline__ float dist3(const float a[3], const float b[3]) {
  const float d0 = __fsub_rn(a[0], b[0]), d1 = __fsub_rn(a[1], b[1]), d2 = __fsub_rn(a[2], b[2]);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
}

__global__ void __launch_bounds__(256) generate_rays_kernel(const __grid_constant__ RayGenArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  int cam;
  float px, py;
  if (a.ray_indices != nullptr) {
    cam = (int)a.ray_indices[i * 3 + 0];
    py = __fadd_rn((float)a.ray_indices[i * 3 + 1], a.pixel_offset);
    px = __fadd_rn((float)a.ray_indices[i * 3 + 2], a.pixel_offset);
  } else {
    cam = a.cam;
    const int64_t p = a.first_pixel + i;
    py = __fadd_rn((float)(p / a.width), a.pixel_offset);
    px = __fadd_rn((float)(p % a.width), a.pixel_offset);
  }
  cam = min(max(cam, 0), a.n_cams - 1);
  const float fx = a.intrinsics[cam * 4 + 0], fy = a.intrinsics[cam * 4 + 1];
  const float cx = a.intrinsics[cam * 4 + 2], cy = a.intrinsics[cam * 4 + 3];
  const float* R = a.c2w + (int64_t)cam * 12;
  const float xc = __fsub_rn(px, cx), yc = __fsub_rn(py, cy);
  const float u = __fdiv_rn(xc, fx), v = -__fdiv_rn(yc, fy);
  const float ux = __fdiv_rn(__fadd_rn(xc, 1.0f), fx), vy = -__fdiv_rn(__fadd_rn(yc, 1.0f), fy);
  float d[3], dxv[3], dyv[3], n;
  rotate_normalize(R, u, v, -1.0f, d, &n);
  rotate_normalize(R, ux, v, -1.0f, dxv, nullptr);
  rotate_normalize(R, u, vy, -1.0f, dyv, nullptr);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a.origins[i * 3 + k] = R[k * 4 + 3];
    a.directions[i * 3 + k] = d[k];
  }
  a.pixel_area[i] = __fmul_rn(dist3(d, dxv), dist3(d, dyv));
  if (a.directions_norm != nullptr) a.directions_norm[i] = n;
  if (a.times != nullptr) a.times[i] = a.cam_times != nullptr ? a.cam_times[cam] : 0.f;
}

}  // namespace kp

using namespace kp;

extern "C" int kp_generate_rays(const float* c2w, const float* intrinsics, const float* cam_times, int n_cams,
                                const int64_t* ray_indices, int cam, int width, int64_t first_pixel, int64_t N,
                                float pixel_offset, float* origins, float* directions, float* pixel_area,
                                float* directions_norm, float* times, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(c2w != nullptr && intrinsics != nullptr && n_cams >= 1, "generate_rays: cameras missing");
  KP_CHECK(origins != nullptr && directions != nullptr && pixel_area != nullptr, "generate_rays: NULL output");
  KP_CHECK(ray_indices != nullptr || (cam >= 0 && cam < n_cams && width >= 1 && first_pixel >= 0),
           "generate_rays: tile mode needs a valid camera (%d of %d), width (%d) and first pixel", cam, n_cams, width);
  KP_CHECK(times == nullptr || cam_times != nullptr, "generate_rays: times requested but the cameras have none");
  RayGenArgs a{c2w, intrinsics, cam_times, ray_indices, n_cams, cam, width, first_pixel, N, pixel_offset,
               origins, directions, pixel_area, directions_norm, times};
  generate_rays_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(a);
  KP_LAUNCH_CHECK("generate_rays");
  return 0;
}
