// Device-resident map maintenance and per-frame scan preparation for the scan-matcher frontend:
// the callers either side of the registration hot path, built so that a frame costs ONE host-to-device copy.
//
// Replaces, in the reference's frontend node (scanmatcher/src/scanmatcher_component.cpp):
//   cloud_callback range filter                      :211-219   -> range_filter_kernel
//   receiveCloud: VoxelGrid(vg_size_for_input) + setInputSource          :323-328   -> b200sm_set_scan
//   initializeMap                                    :257-297   -> b200sm_update_map (first call)
//   updateMap: VoxelGrid(vg_size_for_map), transformPointCloud(Matrix4f), concatenation of the last
//              num_targeted_cloud-1 submaps through transformPointCloud(Affine3d::matrix())      :438-463   -> b200sm_update_map
//   receiveCloud: setInputTarget(targeted) (GICP: VoxelGrid(vg_size_for_input) first)             :300-322   -> b200sm_update_map
//   receiveCloud / publishMapAndPose pose bookkeeping                     :330-353, 391-434     -> b200sm_receive_cloud
// The submaps (sensor-frame, voxel-filtered) and the targeted cloud never leave the GPU; read-back entry points exist for
// the parity tests and for the node's publishers.
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200reg.h"
#include "deskew.hpp"
#include "engine.hpp"

namespace b200 {
namespace {

// r = sqrt(pow(x, 2.0) + pow(y, 2.0)) in double; keep scan_min_range < r < scan_max_range (:213-216). The order of the
// kept points is not preserved (warp-aggregated atomic append): every consumer is a VoxelGrid, which is order-independent.
__global__ void range_filter_kernel(const float4* __restrict__ in, size_t n, double rmin, double rmax, float4* __restrict__ out,
                                    unsigned* __restrict__ count) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  bool keep = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    p = in[i];
    const double r = sqrt((double)p.x * (double)p.x + (double)p.y * (double)p.y);
    keep = (rmin < r) && (r < rmax);
  }
  const unsigned mask = __ballot_sync(0xffffffffu, keep);
  if (mask == 0) return;
  const int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == __ffs(mask) - 1) base = atomicAdd(count, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
  if (keep) out[base + __popc(mask & ((1u << lane) - 1u))] = p;
}

struct Mat34f {
  float m[12];
};
struct Mat34d {
  double m[12];
};

// pcl::transformPointCloud(in, out, Eigen::Matrix4f): xyz <- R xyz + t in float, other fields copied.
// Un-fused, ((m0 x + m1 y) + m2 z) + m3 — the same definition as the solver's transform_point (ndt_solver.cuh).
__global__ void transform_f32_kernel(const float4* __restrict__ in, size_t n, Mat34f T, float4* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  float4 q;
  q.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[0], p.x), __fmul_rn(T.m[1], p.y)), __fmul_rn(T.m[2], p.z)), T.m[3]);
  q.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[4], p.x), __fmul_rn(T.m[5], p.y)), __fmul_rn(T.m[6], p.z)), T.m[7]);
  q.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.m[8], p.x), __fmul_rn(T.m[9], p.y)), __fmul_rn(T.m[10], p.z)), T.m[11]);
  q.w = p.w;
  out[i] = q;
}

// pcl::transformPointCloud(in, out, Eigen::Matrix4d) — the generic Transformer<double>: every coordinate is
// static_cast<float>(m0 x + m1 y + m2 z + m3) evaluated left to right in double (:459-462, submap_affine.matrix()).
__global__ void transform_f64_kernel(const float4* __restrict__ in, size_t n, Mat34d T, float4* __restrict__ out) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const double x = p.x, y = p.y, z = p.z;
  float4 q;
  q.x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[0], x), __dmul_rn(T.m[1], y)), __dmul_rn(T.m[2], z)), T.m[3]);
  q.y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[4], x), __dmul_rn(T.m[5], y)), __dmul_rn(T.m[6], z)), T.m[7]);
  q.z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.m[8], x), __dmul_rn(T.m[9], y)), __dmul_rn(T.m[10], z)), T.m[11]);
  q.w = p.w;
  out[i] = q;
}

// Submap clouds live in a chunked device arena: one cudaMalloc per ~64 MB instead of one per map update (cudaMalloc costs
// tens to hundreds of microseconds and occasionally milliseconds — it would dominate a 0.5 ms frame).
struct SubmapArena {
  static constexpr size_t CHUNK_POINTS = (size_t)4 << 20;
  std::vector<std::unique_ptr<DeviceBuffer<float4>>> chunks;
  size_t used = 0, cap = 0;
  float4* alloc(size_t n) {
    if (chunks.empty() || used + n > cap) {
      chunks.emplace_back(new DeviceBuffer<float4>());
      cap = std::max(CHUNK_POINTS, n);
      chunks.back()->ensure(cap);
      cap = chunks.back()->cap;
      used = 0;
    }
    float4* p = chunks.back()->ptr + used;
    used += (n + 15) & ~(size_t)15;  // keep 256-byte alignment
    return p;
  }
};

struct Submap {
  float4* cloud = nullptr;     // VoxelGrid(vg_size_for_map) of the scan, sensor frame (arena memory)
  size_t n = 0;
  double pose[16];             // row-major 4x4: Translation * Quaternion of the pose the scan was taken at
  double distance = 0;
};

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200reg_adopt_source_device(b200reg_t h, const void* dev, size_t n);  // capi.cu (library-internal)

struct b200sm_session {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // parameters (scanmatcher_component.cpp:34-50 defaults)
  float vg_size_for_input = 0.2f, vg_size_for_map = 0.1f;
  int num_targeted_cloud = 10;
  int use_min_max_filter = 0;
  double scan_min_range = 0.1, scan_max_range = 100.0;
  double trans_for_mapupdate = 1.5;
  // state
  DeviceBuffer<float4> upload, scan;  // uploaded frame; after the optional range filter (`scan` aliases upload when off)
  const float4* d_scan = nullptr;
  size_t n_scan = 0;
  CloudUploader uploader;
  DeviceBuffer<unsigned> counter;
  VoxelGridFilter vg_input, vg_map, vg_target;
  Bounds scan_bounds{};            // min/max of d_scan when it is the uploaded cloud itself (no de-skew, no range filter)
  bool scan_bounds_valid = false;
  size_t n_filtered = 0;
  const float4* d_filtered = nullptr;  // VoxelGrid(vg_size_for_input) of the scan — the scan itself when the grid would overflow
  SubmapArena arena;
  std::vector<std::unique_ptr<Submap>> submaps;
  DeviceBuffer<float4> targeted;
  size_t n_targeted = 0;
  DeviceBuffer<float4> loop_src, loop_tgt;  // search_loop scratch
  int launches = 0;
  // frontend bookkeeping (ScanMatcherComponent members)
  bool initial_cloud_received = false;
  bool target_pending = false;  // is_map_updated_: a rebuilt targeted cloud waits to become the registration target
  double position[3] = {0, 0, 0}, quat[4] = {0, 0, 0, 1};  // corrent_pose_stamped_.pose (x y z, qx qy qz qw)
  double previous_position[3] = {0, 0, 0};
  double latest_distance = 0, trans = 0;
  // IMU de-skew (lidar_undistortion.hpp; use_imu, scanmatcher_component.cpp:205-209)
  ImuDeskew imu;
  bool deskew_armed = false;
  double deskew_scan_time = 0;
};

namespace {

template <typename F>
int sm_guarded(b200sm_t s, F&& f) {
  if (!s) return B200REG_ERR_ARG;
  try {
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) {
      s->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
      return B200REG_ERR_CUDA;
    }
    return f();
  } catch (const CudaError& e) {
    s->err = e.what();
    cudaGetLastError();
    return B200REG_ERR_CUDA;
  } catch (const std::exception& e) {
    s->err = e.what();
    return B200REG_ERR_ARG;
  }
}

int sm_fail(b200sm_t s, int code, const char* msg) {
  s->err = msg;
  return code;
}

// tf2::fromMsg(pose, Affine3d) = Translation3d(p) * Quaterniond(w, x, y, z): Eigen's QuaternionBase::toRotationMatrix
void pose_to_matrix_d(const double* p, const double* q, double* M /* row-major 16 */) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M[0] = 1.0 - (tyy + tzz); M[1] = txy - twz;         M[2] = txz + twy;          M[3] = p[0];
  M[4] = txy + twz;         M[5] = 1.0 - (txx + tzz); M[6] = tyz - twx;          M[7] = p[1];
  M[8] = txz - twy;         M[9] = tyz + twx;         M[10] = 1.0 - (txx + tyy); M[11] = p[2];
  M[12] = 0; M[13] = 0; M[14] = 0; M[15] = 1;
}

// Eigen::Quaterniond(Matrix3d): the trace / largest-diagonal branches of Eigen's quaternionbase_assign_impl (3x3)
void matrix_to_quat_d(const double* R /* row-major 9 */, double* q /* x y z w */) {
  auto m = [&](int r, int c) { return R[r * 3 + c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}

void upload_frame(b200sm_t s, const float* points, size_t n, size_t stride, long intensity_off) {
  s->upload.ensure(n);
  // one H2D copy per frame; the unpack pass also measures the scan's min/max (both VoxelGrids of the frame are sized from it)
  s->uploader.upload_with_bounds(points, n, stride, intensity_off, 0.0f, s->upload.ptr, s->stream);
  s->launches += 1;
  s->d_scan = s->upload.ptr;
  s->n_scan = n;
  s->scan_bounds_valid = false;
  bool deskewed = false;
  if (s->deskew_armed && n > 0) {  // cloud_callback: adjustDistortion before the range filter (sm.cpp:205-209)
    s->deskew_armed = false;
    deskewed = true;
    const char* b = reinterpret_cast<const char*>(points);
    const int before = s->imu.launches;
    s->imu.adjust_distortion(s->upload.ptr, n, reinterpret_cast<const float*>(b), reinterpret_cast<const float*>(b + (n - 1) * stride),
                             s->deskew_scan_time, s->stream);
    s->launches += s->imu.launches - before;
  }
  if (s->use_min_max_filter) {
    s->scan.ensure(n);
    s->counter.ensure(1);
    B200_CUDA(cudaMemsetAsync(s->counter.ptr, 0, sizeof(unsigned), s->stream));
    range_filter_kernel<<<(int)((n + 255) / 256), 256, 0, s->stream>>>(s->upload.ptr, n, s->scan_min_range, s->scan_max_range,
                                                                       s->scan.ptr, s->counter.ptr);
    B200_CUDA(cudaGetLastError());
    unsigned kept = 0;
    B200_CUDA(cudaMemcpyAsync(&kept, s->counter.ptr, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    s->d_scan = s->scan.ptr;
    s->n_scan = kept;
    s->launches += 1;
  } else if (!deskewed) {  // the uploaded cloud IS the scan: its bounds are those measured during the upload
    B200_CUDA(cudaStreamSynchronize(s->stream));
    s->scan_bounds = s->uploader.finish_bounds();
    s->scan_bounds_valid = true;
  }
}

// VoxelGrid on the device; a grid overflow returns the input unchanged like PCL
const float4* filter_on_device(b200sm_t s, VoxelGridFilter& F, const float4* d_in, size_t n, float leaf, size_t* m,
                               const Bounds* known_bounds = nullptr) {
  const int before = F.launches;
  long long cnt = F.filter_device(d_in, n, leaf, s->stream, known_bounds);
  s->launches += F.launches - before;
  if (cnt < 0) {
    *m = n;
    return d_in;
  }
  *m = (size_t)cnt;
  return F.out.ptr;
}

int set_source_from_scan(b200sm_t s, b200reg_t reg) {
  size_t m = 0;
  const float4* f = filter_on_device(s, s->vg_input, s->d_scan, s->n_scan, s->vg_size_for_input, &m,
                                     s->scan_bounds_valid ? &s->scan_bounds : nullptr);
  s->n_filtered = m;
  s->d_filtered = f;
  if (m == 0) return sm_fail(s, B200REG_ERR_ARG, "scan is empty after filtering");
  // (the filter's read-back of the point count has synchronised this stream; the filtered scan lives in this session until
  // the next frame, so the engine reads it in place: no copy, no second synchronisation)
  const int rc = b200reg_adopt_source_device(reg, f, m);
  if (rc != B200REG_OK) s->err = std::string("setInputSource: ") + b200reg_last_error(reg);
  return rc;
}

// updateMap (:438-491) / initializeMap (:257-297) on the current scan; final_T row-major float, pose = position + quaternion
int update_map(b200sm_t s, const float* final_T, const double* position, const double* quat) {
  if (s->n_scan == 0) return sm_fail(s, B200REG_ERR_NO_SOURCE, "update_map: no scan");
  size_t m = 0;
  const float4* filtered = filter_on_device(s, s->vg_map, s->d_scan, s->n_scan, s->vg_size_for_map, &m,
                                            s->scan_bounds_valid ? &s->scan_bounds : nullptr);
  // targeted_cloud_ = T(filtered) + the last num_targeted_cloud-1 submaps, newest first, each through its pose (double)
  const int n_sub = (int)s->submaps.size();
  size_t total = m;
  for (int i = 0; i < s->num_targeted_cloud - 1; i++) {
    if (n_sub - 1 - i < 0) continue;
    total += s->submaps[n_sub - 1 - i]->n;
  }
  s->targeted.ensure(total);
  Mat34f Tf;
  for (int k = 0; k < 12; k++) Tf.m[k] = final_T[k];
  if (m) transform_f32_kernel<<<(int)((m + 255) / 256), 256, 0, s->stream>>>(filtered, m, Tf, s->targeted.ptr);
  size_t off = m;
  for (int i = 0; i < s->num_targeted_cloud - 1; i++) {
    if (n_sub - 1 - i < 0) continue;
    const Submap& sub = *s->submaps[n_sub - 1 - i];
    Mat34d Td;
    for (int k = 0; k < 12; k++) Td.m[k] = sub.pose[k];
    if (sub.n) transform_f64_kernel<<<(int)((sub.n + 255) / 256), 256, 0, s->stream>>>(sub.cloud, sub.n, Td, s->targeted.ptr + off);
    off += sub.n;
    s->launches += 1;
  }
  B200_CUDA(cudaGetLastError());
  s->n_targeted = total;
  s->launches += 1;
  // the new submap keeps the FILTERED, untransformed cloud and the pose (:465-481)
  std::unique_ptr<Submap> sub(new Submap());
  sub->cloud = s->arena.alloc(std::max<size_t>(m, 1));
  sub->n = m;
  if (m) B200_CUDA(cudaMemcpyAsync(sub->cloud, filtered, m * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream));
  pose_to_matrix_d(position, quat, sub->pose);
  sub->distance = s->latest_distance;
  s->submaps.push_back(std::move(sub));
  s->target_pending = true;
  return B200REG_OK;
}

// receiveCloud :300-322: the rebuilt targeted cloud becomes the registration target
int adopt_target(b200sm_t s, b200reg_t reg, int is_gicp) {
  if (!s->target_pending) return B200REG_OK;
  const float4* t = s->targeted.ptr;
  size_t n = s->n_targeted;
  if (is_gicp) t = filter_on_device(s, s->vg_target, s->targeted.ptr, s->n_targeted, s->vg_size_for_input, &n);
  if (n == 0) return sm_fail(s, B200REG_ERR_NO_TARGET, "targeted cloud is empty");
  B200_CUDA(cudaStreamSynchronize(s->stream));
  const int rc = b200reg_set_input_target_device(reg, t, n);
  if (rc != B200REG_OK) {
    s->err = std::string("setInputTarget: ") + b200reg_last_error(reg);
    return rc;
  }
  s->target_pending = false;
  return B200REG_OK;
}

int read_back(b200sm_t s, const float4* d, size_t n, float* out, size_t cap, size_t* n_out) {
  if (n_out) *n_out = n;
  const size_t k = std::min(n, cap);
  if (k && out) {
    B200_CUDA(cudaMemcpyAsync(out, d, k * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
  }
  return B200REG_OK;
}

}  // namespace

extern "C" {

int b200sm_create(int device, b200sm_t* out) {
  if (!out) return B200REG_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return B200REG_ERR_CUDA;  // no CPU fallback
  }
  b200sm_session* s = new b200sm_session();
  s->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete s;
    return B200REG_ERR_CUDA;
  }
  *out = s;
  return B200REG_OK;
}

void b200sm_destroy(b200sm_t s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) {
    cudaStreamSynchronize(s->stream);
    cudaStreamDestroy(s->stream);
  }
  delete s;
}

const char* b200sm_last_error(b200sm_t s) { return s ? s->err.c_str() : "null session"; }

int b200sm_set_params(b200sm_t s, float vg_size_for_input, float vg_size_for_map, int num_targeted_cloud, double trans_for_mapupdate,
                      int use_min_max_filter, double scan_min_range, double scan_max_range) {
  if (!s || !(vg_size_for_input > 0) || !(vg_size_for_map > 0) || num_targeted_cloud < 1) return B200REG_ERR_ARG;
  s->vg_size_for_input = vg_size_for_input;
  s->vg_size_for_map = vg_size_for_map;
  s->num_targeted_cloud = num_targeted_cloud;
  s->trans_for_mapupdate = trans_for_mapupdate;
  s->use_min_max_filter = use_min_max_filter;
  s->scan_min_range = scan_min_range;
  s->scan_max_range = scan_max_range;
  return B200REG_OK;
}

int b200sm_set_initial_pose(b200sm_t s, const double* position3, const double* quat_xyzw) {
  if (!s || !position3 || !quat_xyzw) return B200REG_ERR_ARG;
  for (int k = 0; k < 3; k++) s->position[k] = s->previous_position[k] = position3[k];
  for (int k = 0; k < 4; k++) s->quat[k] = quat_xyzw[k];
  return B200REG_OK;
}

int b200sm_set_scan(b200sm_t s, b200reg_t reg, const float* points, size_t n, size_t stride_bytes, long intensity_offset_bytes,
                    size_t* n_filtered) {
  if (!s || !reg || !points || n == 0 || stride_bytes < 12 || (stride_bytes % 4) != 0 ||
      (intensity_offset_bytes >= 0 && (intensity_offset_bytes % 4) != 0))
    return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    upload_frame(s, points, n, stride_bytes, intensity_offset_bytes);
    const int rc = set_source_from_scan(s, reg);
    if (n_filtered) *n_filtered = s->n_filtered;
    return rc;
  });
}

int b200sm_update_map(b200sm_t s, b200reg_t reg, const float* final_T_colmajor16, const double* position3, const double* quat_xyzw,
                      int adopt_now) {
  if (!s || !final_T_colmajor16 || !position3 || !quat_xyzw) return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    float T[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T[r * 4 + c] = final_T_colmajor16[c * 4 + r];
    if (!s->submaps.empty()) {  // updateMap :471: latest_distance_ += trans_ (distance travelled since the last submap)
      const double dx = position3[0] - s->previous_position[0], dy = position3[1] - s->previous_position[1],
                   dz = position3[2] - s->previous_position[2];
      s->trans = std::sqrt(dx * dx + dy * dy + dz * dz);
      s->latest_distance += s->trans;
    }
    for (int k = 0; k < 3; k++) s->previous_position[k] = position3[k];
    int rc = update_map(s, T, position3, quat_xyzw);
    if (rc == B200REG_OK && adopt_now && reg) {
      int kind = B200REG_NDT;
      b200reg_get_kind(reg, &kind);
      rc = adopt_target(s, reg, kind == B200REG_GICP);
    }
    return rc;
  });
}

int b200sm_receive_cloud(b200sm_t s, b200reg_t reg, const float* points, size_t n, size_t stride_bytes, long intensity_offset_bytes,
                         double* pose7_out, float* final_T_colmajor16_out, int* map_updated) {
  if (!s || !reg || !points || n == 0 || stride_bytes < 12 || (stride_bytes % 4) != 0 ||
      (intensity_offset_bytes >= 0 && (intensity_offset_bytes % 4) != 0))
    return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    if (map_updated) *map_updated = 0;
    int kind = B200REG_NDT;
    b200reg_get_kind(reg, &kind);
    upload_frame(s, points, n, stride_bytes, intensity_offset_bytes);
    // sim_trans = getTransformation(corrent_pose_stamped_.pose): Affine3d matrix cast to float (:493-499)
    double M[16];
    float sim_col[16], T_row[16];
    auto sim_trans = [&]() {
      pose_to_matrix_d(s->position, s->quat, M);
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          T_row[r * 4 + c] = (float)M[r * 4 + c];
          sim_col[c * 4 + r] = (float)M[r * 4 + c];
        }
    };
    int rc;
    if (!s->initial_cloud_received) {  // initializeMap (:257-297): the first scan, at the initial pose, is the map
      s->initial_cloud_received = true;
      sim_trans();
      rc = update_map(s, T_row, s->position, s->quat);
      if (rc != B200REG_OK) return rc;
      rc = adopt_target(s, reg, /*is_gicp=*/0);  // initializeMap hands the transformed cloud over unfiltered
      if (rc != B200REG_OK) return rc;
    }
    rc = adopt_target(s, reg, kind == B200REG_GICP);  // :300-322
    if (rc != B200REG_OK) return rc;
    rc = set_source_from_scan(s, reg);                // :323-328
    if (rc != B200REG_OK) return rc;
    sim_trans();
    float final_col[16];
    rc = b200reg_align(reg, sim_col, final_col);      // :350
    if (rc != B200REG_OK) {
      s->err = std::string("align: ") + b200reg_last_error(reg);
      return rc;
    }
    // publishMapAndPose (:391-434)
    double R[9], pos[3];
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) R[r * 3 + c] = (double)final_col[c * 4 + r];
      pos[r] = (double)final_col[12 + r];
    }
    matrix_to_quat_d(R, s->quat);
    for (int k = 0; k < 3; k++) s->position[k] = pos[k];
    const double dx = pos[0] - s->previous_position[0], dy = pos[1] - s->previous_position[1], dz = pos[2] - s->previous_position[2];
    s->trans = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (s->trans >= s->trans_for_mapupdate) {
      for (int k = 0; k < 3; k++) s->previous_position[k] = pos[k];
      s->latest_distance += s->trans;  // updateMap :471
      float F_row[16];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) F_row[r * 4 + c] = final_col[c * 4 + r];
      rc = update_map(s, F_row, s->position, s->quat);
      if (rc != B200REG_OK) return rc;
      if (map_updated) *map_updated = 1;
    }
    if (pose7_out) {
      for (int k = 0; k < 3; k++) pose7_out[k] = s->position[k];
      for (int k = 0; k < 4; k++) pose7_out[3 + k] = s->quat[k];
    }
    if (final_T_colmajor16_out) std::memcpy(final_T_colmajor16_out, final_col, sizeof(final_col));
    return (int)B200REG_OK;
  });
}

int b200sm_num_submaps(b200sm_t s, size_t* out) {
  if (!s || !out) return B200REG_ERR_ARG;
  *out = s->submaps.size();
  return B200REG_OK;
}

int b200sm_get_targeted(b200sm_t s, float* out_xyzi, size_t capacity, size_t* n) {
  if (!s) return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() { return read_back(s, s->targeted.ptr, s->n_targeted, out_xyzi, capacity, n); });
}

int b200sm_get_submap(b200sm_t s, size_t index, float* out_xyzi, size_t capacity, size_t* n, double* pose_colmajor16, double* distance) {
  if (!s || index >= s->submaps.size()) return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    const Submap& sub = *s->submaps[index];
    if (pose_colmajor16)
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) pose_colmajor16[c * 4 + r] = sub.pose[r * 4 + c];
    if (distance) *distance = sub.distance;
    return read_back(s, sub.cloud, sub.n, out_xyzi, capacity, n);
  });
}

int b200sm_get_filtered_scan(b200sm_t s, float* out_xyzi, size_t capacity, size_t* n) {
  if (!s) return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() { return read_back(s, s->d_filtered, s->d_filtered ? s->n_filtered : 0, out_xyzi, capacity, n); });
}

// GraphBasedSlamComponent::searchLoop (graph_based_slam_component.cpp:144-258) over the session's own submaps — the map
// array the frontend publishes is the backend's input, here it never left the device.
}  // extern "C"

namespace {

Mat34f pose_f32(const Submap& sub) {  // affine.matrix().cast<float>()
  Mat34f T;
  for (int k = 0; k < 12; k++) T.m[k] = (float)sub.pose[k];
  return T;
}

struct LoopCandidate {
  int id;
  double dist;
};

// the gates of :187-204: travelled distance apart, position close — every submap that passes them, ascending id
std::vector<LoopCandidate> loop_candidates(b200sm_t s, double distance_loop_closure, double range_of_searching_loop_closure) {
  std::vector<LoopCandidate> out;
  const int n_sub = (int)s->submaps.size();
  if (n_sub == 0) return out;
  const Submap& latest = *s->submaps[n_sub - 1];
  for (int i = 0; i < n_sub; i++) {
    const Submap& sub = *s->submaps[i];
    const double dx = latest.pose[3] - sub.pose[3], dy = latest.pose[7] - sub.pose[7], dz = latest.pose[11] - sub.pose[11];
    const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    if (latest.distance - sub.distance > distance_loop_closure && dist < range_of_searching_loop_closure) out.push_back({i, dist});
  }
  return out;
}

// source = latest submap in the map frame (:165-176), handed to the registration object once per search
int loop_set_source(b200sm_t s, b200reg_t reg) {
  const Submap& latest = *s->submaps.back();
  s->loop_src.ensure(std::max<size_t>(latest.n, 1));
  if (latest.n == 0) return sm_fail(s, B200REG_ERR_NO_SOURCE, "search_loop: empty source");
  transform_f32_kernel<<<(int)((latest.n + 255) / 256), 256, 0, s->stream>>>(latest.cloud, latest.n, pose_f32(latest), s->loop_src.ptr);
  B200_CUDA(cudaGetLastError());
  s->launches += 1;
  B200_CUDA(cudaStreamSynchronize(s->stream));
  const int rc = b200reg_set_input_source_device(reg, s->loop_src.ptr, latest.n);
  if (rc != B200REG_OK) s->err = std::string("search_loop: ") + b200reg_last_error(reg);
  return rc;
}

// one candidate: target = VoxelGrid(voxel_leaf_size) of the submaps id - search_submap_num .. id + search_submap_num
// (:206-225), align without guess (:229), getFitnessScore (:230), loop edge when the score passes (:232-246).
// The reference does not test the upper index (undefined behaviour when the window runs past the newest submap);
// here indices beyond the array are skipped like the negative ones. Nothing is uploaded: the submaps live in HBM.
int loop_evaluate(b200sm_t s, b200reg_t reg, const LoopCandidate& cand, float voxel_leaf_size, double threshold_loop_closure_score,
                  int search_submap_num, b200sm_loop_result* out) {
  const int n_sub = (int)s->submaps.size();
  const Submap& latest = *s->submaps[n_sub - 1];
  const int id_min = cand.id;
  out->is_candidate = 1;
  out->id_min = id_min;
  out->min_dist = cand.dist;
  size_t total = 0;
  for (int j = 0; j <= 2 * search_submap_num; j++) {
    const int idx = id_min + j - search_submap_num;
    if (idx < 0 || idx >= n_sub) continue;
    total += s->submaps[idx]->n;
  }
  s->loop_tgt.ensure(std::max<size_t>(total, 1));
  size_t off = 0;
  for (int j = 0; j <= 2 * search_submap_num; j++) {
    const int idx = id_min + j - search_submap_num;
    if (idx < 0 || idx >= n_sub) continue;
    const Submap& sub = *s->submaps[idx];
    if (sub.n) transform_f32_kernel<<<(int)((sub.n + 255) / 256), 256, 0, s->stream>>>(sub.cloud, sub.n, pose_f32(sub), s->loop_tgt.ptr + off);
    off += sub.n;
    s->launches += 1;
  }
  B200_CUDA(cudaGetLastError());
  size_t m = 0;
  const float4* tgt = filter_on_device(s, s->vg_target, s->loop_tgt.ptr, total, voxel_leaf_size, &m);
  if (m == 0) return sm_fail(s, B200REG_ERR_NO_TARGET, "search_loop: empty target");
  B200_CUDA(cudaStreamSynchronize(s->stream));
  int rc = b200reg_set_input_target_device(reg, tgt, m);
  float fin[16];
  if (rc == B200REG_OK) rc = b200reg_align(reg, nullptr, fin);  // :229, no guess
  double fitness = 0;
  if (rc == B200REG_OK) rc = b200reg_get_fitness_score(reg, 1.7976931348623157e308, &fitness);  // :230
  if (rc != B200REG_OK) {
    s->err = std::string("search_loop: ") + b200reg_last_error(reg);
    return rc;
  }
  out->n_source = latest.n;
  out->n_target = m;
  out->fitness = fitness;
  std::memcpy(out->final_T, fin, sizeof(fin));
  if (fitness < threshold_loop_closure_score) {  // :232-246: loop edge (id_min, newest), relative pose from^-1 * (final * init)
    out->accepted = 1;
    double F[16], to[16], rel[16];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) F[r * 4 + c] = (double)fin[c * 4 + r];
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        double a = 0;
        for (int k = 0; k < 4; k++) a += F[r * 4 + k] * latest.pose[k * 4 + c];
        to[r * 4 + c] = a;
      }
    const double* fr = s->submaps[id_min]->pose;  // Isometry3d::inverse(): R^T, -R^T t
    double inv[16] = {fr[0], fr[4], fr[8], 0, fr[1], fr[5], fr[9], 0, fr[2], fr[6], fr[10], 0, 0, 0, 0, 1};
    for (int r = 0; r < 3; r++) inv[r * 4 + 3] = -(inv[r * 4 + 0] * fr[3] + inv[r * 4 + 1] * fr[7] + inv[r * 4 + 2] * fr[11]);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        double a = 0;
        for (int k = 0; k < 4; k++) a += inv[r * 4 + k] * to[k * 4 + c];
        rel[r * 4 + c] = a;
      }
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) out->relative_pose[c * 4 + r] = rel[r * 4 + c];
  }
  return B200REG_OK;
}

}  // namespace

extern "C" {

int b200sm_search_loop(b200sm_t s, b200reg_t reg, float voxel_leaf_size, double threshold_loop_closure_score,
                       double distance_loop_closure, double range_of_searching_loop_closure, int search_submap_num,
                       b200sm_loop_result* out) {
  if (!s || !reg || !out || !(voxel_leaf_size > 0) || search_submap_num < 0) return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    std::memset(out, 0, sizeof(*out));
    out->id_min = -1;
    // the closest of the submaps that pass the gates wins (:193-201; the first one on a tie, like the strict `<`)
    const std::vector<LoopCandidate> cands = loop_candidates(s, distance_loop_closure, range_of_searching_loop_closure);
    if (cands.empty()) return (int)B200REG_OK;
    LoopCandidate best = cands[0];
    for (const LoopCandidate& c : cands)
      if (c.dist < best.dist) best = c;
    out->is_candidate = 1;
    out->id_min = best.id;
    out->min_dist = best.dist;
    int rc = loop_set_source(s, reg);
    if (rc != B200REG_OK) return rc;
    return loop_evaluate(s, reg, best, voxel_leaf_size, threshold_loop_closure_score, search_submap_num, out);
  });
}

// The generalisation SURVEY.md section 8f row 2 names: EVERY submap that passes the two gates is registered against the
// newest one (the reference keeps only the closest), all on the device-resident submaps. shard_rank / shard_world
// (0 / 1 on one GPU) deal the candidates out across processes: candidate k (ascending submap id) belongs to rank
// k mod shard_world; the caller all-gathers the rows (include/b200comm.h).
int b200sm_search_loop_all(b200sm_t s, b200reg_t reg, float voxel_leaf_size, double threshold_loop_closure_score,
                           double distance_loop_closure, double range_of_searching_loop_closure, int search_submap_num,
                           int shard_rank, int shard_world, b200sm_loop_result* out, size_t capacity, size_t* n_out,
                           size_t* n_candidates_total) {
  if (!s || !reg || !n_out || !(voxel_leaf_size > 0) || search_submap_num < 0 || shard_world < 1 || shard_rank < 0 ||
      shard_rank >= shard_world || (!out && capacity))
    return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    *n_out = 0;
    const std::vector<LoopCandidate> cands = loop_candidates(s, distance_loop_closure, range_of_searching_loop_closure);
    if (n_candidates_total) *n_candidates_total = cands.size();
    if (cands.empty()) return (int)B200REG_OK;
    bool have_source = false;
    for (size_t k = 0; k < cands.size(); k++) {
      if ((int)(k % (size_t)shard_world) != shard_rank) continue;
      if (*n_out >= capacity) break;
      if (!have_source) {
        const int rc = loop_set_source(s, reg);
        if (rc != B200REG_OK) return rc;
        have_source = true;
      }
      b200sm_loop_result* r = out + *n_out;
      std::memset(r, 0, sizeof(*r));
      const int rc = loop_evaluate(s, reg, cands[k], voxel_leaf_size, threshold_loop_closure_score, search_submap_num, r);
      if (rc != B200REG_OK) return rc;
      *n_out += 1;
    }
    return (int)B200REG_OK;
  });
}

// A backend that runs in its own process receives the frontend's submaps as lidarslam_msgs/SubMap messages (already
// voxel-filtered cloud in the sensor frame + pose + travelled distance, graph_based_slam_component.cpp:91-101): this entry
// point appends one to the session so that b200sm_search_loop / _all work on device-resident copies there too.
int b200sm_import_submap(b200sm_t s, const float* points, size_t n, size_t stride_bytes, long intensity_offset_bytes,
                         const double* pose_colmajor16, double distance) {
  if (!s || (!points && n) || !pose_colmajor16 || stride_bytes < 12 || (stride_bytes % 4) != 0 ||
      (intensity_offset_bytes >= 0 && intensity_offset_bytes % 4 != 0))
    return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    std::unique_ptr<Submap> sub(new Submap());
    sub->cloud = s->arena.alloc(std::max<size_t>(n, 1));
    sub->n = n;
    if (n) {
      s->uploader.upload(points, n, stride_bytes, intensity_offset_bytes, 0.0f, sub->cloud, s->stream);
      B200_CUDA(cudaStreamSynchronize(s->stream));  // the caller may reuse its buffer
      s->launches += 1;
    }
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) sub->pose[r * 4 + c] = pose_colmajor16[c * 4 + r];
    sub->distance = distance;
    s->latest_distance = distance;
    s->submaps.push_back(std::move(sub));
    return (int)B200REG_OK;
  });
}

int b200sm_get_stats(b200sm_t s, b200sm_stats* out) {
  if (!s || !out) return B200REG_ERR_ARG;
  out->n_scan = s->n_scan;
  out->n_filtered = s->n_filtered;
  out->n_targeted = s->n_targeted;
  out->n_submaps = s->submaps.size();
  out->kernel_launches = s->launches;
  out->trans = s->trans;
  out->latest_distance = s->latest_distance;
  return B200REG_OK;
}

}  // extern "C"

// ---- IMU de-skew: LidarUndistortion of scanmatcher/include/scanmatcher/lidar_undistortion.hpp ----
extern "C" {

int b200sm_imu_set_scan_period(b200sm_t s, double scan_period) {
  if (!s || !(scan_period > 0)) return B200REG_ERR_ARG;
  s->imu.scan_period = scan_period;
  return B200REG_OK;
}

int b200sm_imu_push(b200sm_t s, const float* angular_velocity3, const float* linear_acceleration3, const float* orientation_xyzw,
                    double stamp) {
  if (!s || !angular_velocity3 || !linear_acceleration3 || !orientation_xyzw) return B200REG_ERR_ARG;
  s->imu.get_imu(angular_velocity3, linear_acceleration3, orientation_xyzw, stamp);
  return B200REG_OK;
}

int b200sm_deskew_next_scan(b200sm_t s, double scan_time) {
  if (!s) return B200REG_ERR_ARG;
  s->deskew_armed = true;
  s->deskew_scan_time = scan_time;
  return B200REG_OK;
}

int b200sm_imu_adjust_distortion(b200sm_t s, float* points, size_t n, size_t stride_bytes, long intensity_offset_bytes,
                                 double scan_time) {
  if (!s || (!points && n) || stride_bytes < 12 || (stride_bytes % 4) != 0 || (intensity_offset_bytes >= 0 && intensity_offset_bytes % 4 != 0))
    return B200REG_ERR_ARG;
  return sm_guarded(s, [&]() {
    if (n == 0) return (int)B200REG_OK;
    s->upload.ensure(n);
    s->uploader.upload(points, n, stride_bytes, intensity_offset_bytes, 0.0f, s->upload.ptr, s->stream);
    char* b = reinterpret_cast<char*>(points);
    const int before = s->imu.launches;
    s->imu.adjust_distortion(s->upload.ptr, n, reinterpret_cast<const float*>(b), reinterpret_cast<const float*>(b + (n - 1) * stride_bytes),
                             scan_time, s->stream);
    s->launches += 1 + s->imu.launches - before;
    std::vector<float4> host(n);
    B200_CUDA(cudaMemcpyAsync(host.data(), s->upload.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t i = 0; i < n; i++) {  // x, y, z back into the caller's records; every other field is untouched
      float* f = reinterpret_cast<float*>(b + i * stride_bytes);
      f[0] = host[i].x;
      f[1] = host[i].y;
      f[2] = host[i].z;
    }
    return (int)B200REG_OK;
  });
}

int b200sm_imu_get_state(b200sm_t s, int* ptr_front, int* ptr_last, int* ptr_last_iter) {
  if (!s) return B200REG_ERR_ARG;
  if (ptr_front) *ptr_front = s->imu.ptr_front;
  if (ptr_last) *ptr_last = s->imu.ptr_last;
  if (ptr_last_iter) *ptr_last_iter = s->imu.ptr_last_iter;
  return B200REG_OK;
}

int b200sm_imu_get_sample(b200sm_t s, int index, double* stamp, float* rpy3, float* shift3, float* velo3) {
  if (!s || index < 0 || index >= IMU_QUE) return B200REG_ERR_ARG;
  if (stamp) *stamp = s->imu.time[index];
  if (rpy3) {
    rpy3[0] = s->imu.roll[index];
    rpy3[1] = s->imu.pitch[index];
    rpy3[2] = s->imu.yaw[index];
  }
  for (int c = 0; c < 3; c++) {
    if (shift3) shift3[c] = s->imu.shift[index][c];
    if (velo3) velo3[c] = s->imu.velo[index][c];
  }
  return B200REG_OK;
}

}  // extern "C"
