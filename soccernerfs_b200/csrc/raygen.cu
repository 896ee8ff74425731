// Pixel -> ray generation on the device (sm_100a): Cameras._generate_rays_from_coords (NS/cameras/cameras.py:505-741)
// for perspective, fisheye and equirectangular cameras with OpenCV lens distortion, as used by RayGenerator.forward
// (NS/model_components/ray_generators.py:43-59, training: explicit (camera,row,col) triplets) and by
// Cameras.generate_rays(camera_indices=i, keep_shape=True) (cameras.py:327-502, evaluation: a row-major tile of a
// full frame).  One thread per ray; the reference's operation order is kept (no FMA contraction): image-plane
// coordinates ((x - cx) / fx, -(y - cy) / fy) of the pixel centre and of its +1 neighbours in x and y, rotation by
// the camera-to-world matrix as three products summed left to right, normalisation by max(norm, 4*DBL_EPSILON),
// pixel_area = |d - d_x| * |d - d_y|.  A batch of undistorted perspective cameras takes the <false> instantiation
// (the arithmetic above and nothing else); camera types and distortion parameters select <true>: the three image-plane
// points are undistorted by the reference's 10 Newton iterations (camera_utils.py:298-401, same expression order) unless
// the camera is equirectangular (cameras.py:645-654), then mapped to a camera-space direction per camera type
// (cameras.py:665-697).
#include "common.cuh"

namespace kp {

struct RayGenArgs {
  const float* c2w;         // [n_cams,3,4]
  const float* intrinsics;  // [n_cams,4] = fx, fy, cx, cy
  const float* cam_times;   // [n_cams] or null
  const int64_t* ray_indices;  // [N,3] = camera,row,col or null (tile mode)
  const float* distortion;  // [n_cams,6] = k1,k2,k3,k4,p1,p2 or null
  const int32_t* cam_types; // [n_cams] CameraType values (1 perspective, 2 fisheye, 3 equirectangular) or null
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

// radial_and_tangential_undistort (camera_utils.py:363-401) with _compute_residual_and_jacobian (:298-360): Newton's
// method on the OpenCV forward model from the distorted point, max_iterations = 10, a step only where |det| > eps = 1e-3.
// Every product / sum is rounded on its own, in the order the reference's tensor expressions evaluate.
__device__ __forceinline__ void undistort_point(const float* __restrict__ k, float xd, float yd, float* xo, float* yo) {
  const float k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3], p1 = k[4], p2 = k[5];
  const float p1_2 = __fmul_rn(2.0f, p1), p2_2 = __fmul_rn(2.0f, p2), p1_6 = __fmul_rn(6.0f, p1), p2_6 = __fmul_rn(6.0f, p2);
  const float k2_2 = __fmul_rn(2.0f, k2), k3_3 = __fmul_rn(3.0f, k3);
  float x = xd, y = yd;
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    const float r = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float d = __fadd_rn(1.0f, __fmul_rn(r, __fadd_rn(k1, __fmul_rn(r, __fadd_rn(k2, __fmul_rn(r, __fadd_rn(k3, __fmul_rn(r, k4))))))));
    const float x2 = __fmul_rn(2.0f, x), y2 = __fmul_rn(2.0f, y);
    // fx = d*x + 2*p1*x*y + p2*(r + 2*x*x) - xd ;  fy = d*y + 2*p2*x*y + p1*(r + 2*y*y) - yd
    const float fx = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(d, x), __fmul_rn(__fmul_rn(p1_2, x), y)),
                                         __fmul_rn(p2, __fadd_rn(r, __fmul_rn(x2, x)))), xd);
    const float fy = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(d, y), __fmul_rn(__fmul_rn(p2_2, x), y)),
                                         __fmul_rn(p1, __fadd_rn(r, __fmul_rn(y2, y)))), yd);
    // d_r = k1 + r*(2*k2 + r*(3*k3 + r*4*k4))
    const float d_r = __fadd_rn(k1, __fmul_rn(r, __fadd_rn(k2_2, __fmul_rn(r, __fadd_rn(k3_3, __fmul_rn(__fmul_rn(r, 4.0f), k4))))));
    const float d_x = __fmul_rn(x2, d_r), d_y = __fmul_rn(y2, d_r);
    const float fx_x = __fadd_rn(__fadd_rn(__fadd_rn(d, __fmul_rn(d_x, x)), __fmul_rn(p1_2, y)), __fmul_rn(p2_6, x));
    const float fx_y = __fadd_rn(__fadd_rn(__fmul_rn(d_y, x), __fmul_rn(p1_2, x)), __fmul_rn(p2_2, y));
    const float fy_x = __fadd_rn(__fadd_rn(__fmul_rn(d_x, y), __fmul_rn(p2_2, y)), __fmul_rn(p1_2, x));
    const float fy_y = __fadd_rn(__fadd_rn(__fadd_rn(d, __fmul_rn(d_y, y)), __fmul_rn(p2_2, x)), __fmul_rn(p1_6, y));
    const float den = __fsub_rn(__fmul_rn(fy_x, fx_y), __fmul_rn(fx_x, fy_y));
    const float xn = __fsub_rn(__fmul_rn(fx, fy_y), __fmul_rn(fy, fx_y));
    const float yn = __fsub_rn(__fmul_rn(fy, fx_x), __fmul_rn(fx, fy_x));
    const bool ok = fabsf(den) > 1e-3f;
    x = __fadd_rn(x, ok ? __fdiv_rn(xn, den) : 0.0f);
    y = __fadd_rn(y, ok ? __fdiv_rn(yn, den) : 0.0f);
  }
  *xo = x;
  *yo = y;
}

// image-plane point -> camera-space direction (before the rotation), cameras.py:665-697
__device__ __forceinline__ void lens_direction(int type, const float* __restrict__ k, float u, float v, float out[3]) {
  if (k != nullptr && type != 3) undistort_point(k, u, v, &u, &v);
  if (type == 2) {  // fisheye: equidistant model, theta = |coord| clipped to [0, pi]
    float theta = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
    theta = fminf(fmaxf(theta, 0.0f), 3.14159274101257324f);
    const float st = sinf(theta);
    out[0] = __fdiv_rn(__fmul_rn(u, st), theta);
    out[1] = __fdiv_rn(__fmul_rn(v, st), theta);
    out[2] = -cosf(theta);
  } else if (type == 3) {  // equirectangular: longitude from u, latitude from v
    const float theta = __fmul_rn(-3.14159274101257324f, u);
    const float phi = __fmul_rn(3.14159274101257324f, __fsub_rn(0.5f, v));
    const float sp = sinf(phi);
    out[0] = __fmul_rn(-sinf(theta), sp);
    out[1] = cosf(phi);
    out[2] = __fmul_rn(-cosf(theta), sp);
  } else {
    out[0] = u;
    out[1] = v;
    out[2] = -1.0f;
  }
}

template <bool LENS>
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
  if constexpr (LENS) {
    const int type = a.cam_types != nullptr ? a.cam_types[cam] : 1;
    const float* k = a.distortion != nullptr ? a.distortion + (int64_t)cam * 6 : nullptr;
    float c0[3], c1[3], c2[3];
    lens_direction(type, k, u, v, c0);
    lens_direction(type, k, ux, v, c1);
    lens_direction(type, k, u, vy, c2);
    rotate_normalize(R, c0[0], c0[1], c0[2], d, &n);
    rotate_normalize(R, c1[0], c1[1], c1[2], dxv, nullptr);
    rotate_normalize(R, c2[0], c2[1], c2[2], dyv, nullptr);
  } else {
    rotate_normalize(R, u, v, -1.0f, d, &n);
    rotate_normalize(R, ux, v, -1.0f, dxv, nullptr);
    rotate_normalize(R, u, vy, -1.0f, dyv, nullptr);
  }
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

extern "C" int kp_generate_rays(const float* c2w, const float* intrinsics, const float* cam_times,
                                const float* distortion, const int32_t* cam_types, int n_cams,
                                const int64_t* ray_indices, int cam, int width, int64_t first_pixel, int64_t N,
                                float pixel_offset, float* origins, float* directions, float* pixel_area,
                                float* directions_norm, float* times, void* stream) {
  if (N == 0) return 0;
  KP_CHECK(c2w != nullptr && intrinsics != nullptr && n_cams >= 1, "generate_rays: cameras missing");
  KP_CHECK(origins != nullptr && directions != nullptr && pixel_area != nullptr, "generate_rays: NULL output");
  KP_CHECK(ray_indices != nullptr || (cam >= 0 && cam < n_cams && width >= 1 && first_pixel >= 0),
           "generate_rays: tile mode needs a valid camera (%d of %d), width (%d) and first pixel", cam, n_cams, width);
  KP_CHECK(times == nullptr || cam_times != nullptr, "generate_rays: times requested but the cameras have none");
  RayGenArgs a{c2w, intrinsics, cam_times, ray_indices, distortion, cam_types, n_cams, cam, width, first_pixel, N,
               pixel_offset, origins, directions, pixel_area, directions_norm, times};
  if (distortion != nullptr || cam_types != nullptr)
    generate_rays_kernel<true><<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(a);
  else
    generate_rays_kernel<false><<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(a);
  KP_LAUNCH_CHECK("generate_rays");
  return 0;
}
