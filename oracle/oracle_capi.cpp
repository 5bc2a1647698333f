// oracle_capi.cpp -- TEST INFRASTRUCTURE (see vloam_oracle.h header).
// Plain C entry points so tests/ and bench.py's cpu_baseline leg can drive the
// oracle through ctypes.  Mirrors the name-keyed debug interface of the
// product library (include/vloam_b200.h) so parity tests compare like with like.
#include <string.h>
#include <string>
#include <vector>
#include "vloam_oracle.h"

using namespace vo;

struct OracleParamsC {  // same layout as vloam_b200_params
  int n_scans;
  float minimum_range;
  float line_res, plane_res;
  int mapping_skip_frame;
  int knn_backend;
  int distortion;  // LaserOdometry::DISTORTION
};

static int put(const void* src, size_t bytes, void* out, long cap) {
  if (out && (long)bytes <= cap && bytes) memcpy(out, src, bytes);
  return (int)bytes;
}
template <typename T> static int putv(const std::vector<T>& v, void* out, long cap) { return put(v.data(), v.size() * sizeof(T), out, cap); }

static int map_export(const std::vector<Cloud*>& arr, void* out, long cap) {
  std::vector<char> blob(LaserMapping::NUM * 4);
  int* counts = (int*)blob.data();
  size_t total = 0;
  for (int i = 0; i < LaserMapping::NUM; ++i) { counts[i] = (int)arr[i]->size(); total += arr[i]->size(); }
  blob.resize(LaserMapping::NUM * 4 + total * 16);
  char* p = blob.data() + LaserMapping::NUM * 4;
  for (int i = 0; i < LaserMapping::NUM; ++i) { memcpy(p, arr[i]->data(), arr[i]->size() * 16); p += arr[i]->size() * 16; }
  return put(blob.data(), blob.size(), out, cap);
}
static void map_import(std::vector<Cloud*>& arr, const void* data) {
  const int* counts = (const int*)data;
  const P4* p = (const P4*)((const char*)data + LaserMapping::NUM * 4);
  for (int i = 0; i < LaserMapping::NUM; ++i) { arr[i]->assign(p, p + counts[i]); p += counts[i]; }
}

extern "C" {

void* vloam_oracle_create(const OracleParamsC* pc) {
  Params p;
  p.n_scans = pc->n_scans; p.minimum_range = pc->minimum_range; p.line_res = pc->line_res; p.plane_res = pc->plane_res;
  p.mapping_skip_frame = pc->mapping_skip_frame; p.knn_backend = pc->knn_backend; p.distortion = pc->distortion;
  return new Pipeline(p);
}
void vloam_oracle_destroy(void* h) { delete (Pipeline*)h; }

// One frame through SR -> LO -> LM (MAIN.cpp:143-144, 186-190).
int vloam_oracle_process(void* h, const float* xyz, int n, int stride) { ((Pipeline*)h)->process(xyz, n, stride); return 0; }

// Stage-by-stage driving (teacher-forced tests).
int vloam_oracle_scan_registration(void* h, const float* xyz, int n, int stride) {
  Pipeline* p = (Pipeline*)h; p->sr.reset(); p->lm.reset(); p->sr.input(xyz, n, stride); return 0;
}
int vloam_oracle_laser_odometry(void* h, const double* prior_q, const double* prior_t, int use_prior) {
  Pipeline* p = (Pipeline*)h;
  p->lo.input(p->sr.laserCloud, p->sr.cornerSharp, p->sr.cornerLessSharp, p->sr.surfFlat, p->sr.surfLessFlat);
  p->lo.solveLO(prior_q, prior_t, use_prior != 0);
  return p->lo.skip_frame() ? 1 : 0;
}
int vloam_oracle_laser_mapping(void* h) {
  Pipeline* p = (Pipeline*)h;
  const bool skip = p->lo.skip_frame();
  p->lm.input(p->lo.cornerLast, p->lo.surfLast, p->lo.q_w, p->lo.t_w, skip);
  if (!skip) p->lm.solveMapping();
  return 0;
}
// LO association only at pose x = {qx,qy,qz,qw,tx,ty,tz}; current sharp/flat vs last clouds.
int vloam_oracle_lo_associate(void* h, const double* x, int* corner_idx, int* surf_idx) {
  Pipeline* p = (Pipeline*)h;
  p->lo.input(p->sr.laserCloud, p->sr.cornerSharp, p->sr.cornerLessSharp, p->sr.surfFlat, p->sr.surfLessFlat);
  std::vector<int> ci, si;
  p->lo.associate(x, x + 4, nullptr, &ci, &si);
  if (corner_idx) memcpy(corner_idx, ci.data(), ci.size() * 4);
  if (surf_idx) memcpy(surf_idx, si.data(), si.size() * 4);
  return 0;
}

int vloam_oracle_get(void* h, const char* name, void* out, long cap) {
  Pipeline* p = (Pipeline*)h;
  const std::string n(name);
  if (n == "sr.laserCloud") return putv(p->sr.laserCloud, out, cap);
  if (n == "sr.sharp") return putv(p->sr.cornerSharp, out, cap);
  if (n == "sr.lessSharp") return putv(p->sr.cornerLessSharp, out, cap);
  if (n == "sr.flat") return putv(p->sr.surfFlat, out, cap);
  if (n == "sr.lessFlat") return putv(p->sr.surfLessFlat, out, cap);
  if (n == "sr.curvature") return putv(p->sr.curvature, out, cap);
  if (n == "sr.label") return putv(p->sr.label, out, cap);
  if (n == "sr.scanStartInd") return putv(p->sr.scanStartInd, out, cap);
  if (n == "sr.scanEndInd") return putv(p->sr.scanEndInd, out, cap);
  if (n == "lo.cornerLast") return putv(p->lo.cornerLast, out, cap);
  if (n == "lo.surfLast") return putv(p->lo.surfLast, out, cap);
  if (n == "lo.pose") {  // q_w[4] t_w[3] para_q[4] para_t[3]
    double v[14]; memcpy(v, p->lo.q_w, 32); memcpy(v + 4, p->lo.t_w, 24); memcpy(v + 7, p->lo.para_q, 32); memcpy(v + 11, p->lo.para_t, 24);
    return put(v, sizeof v, out, cap);
  }
  if (n == "lo.assoc.corner0") return putv(p->lo.dbg_corner[0], out, cap);
  if (n == "lo.assoc.corner1") return putv(p->lo.dbg_corner[1], out, cap);
  if (n == "lo.assoc.surf0") return putv(p->lo.dbg_surf[0], out, cap);
  if (n == "lo.assoc.surf1") return putv(p->lo.dbg_surf[1], out, cap);
  if (n == "lo.costs") {  // initial/final cost of the two passes
    double v[4] = {p->lo.dbg_log[0].initial_cost, p->lo.dbg_log[0].final_cost, p->lo.dbg_log[1].initial_cost, p->lo.dbg_log[1].final_cost};
    return put(v, sizeof v, out, cap);
  }
  if (n == "lm.pose") {  // q_w_curr[4] t_w_curr[3] q_wmap_wodom[4] t_wmap_wodom[3]
    double v[14]; memcpy(v, p->lm.parameters, 56); memcpy(v + 7, p->lm.q_wmap_wodom, 32); memcpy(v + 11, p->lm.t_wmap_wodom, 24);
    return put(v, sizeof v, out, cap);
  }
  if (n == "lm.poseHighFreq") {  // q_w_curr_highfreq[4] t_w_curr_highfreq[3] (LM.cpp:197-201: the pose of a skipped frame)
    double v[7]; memcpy(v, p->lm.q_hf, 32); memcpy(v + 4, p->lm.t_hf, 24);
    return put(v, sizeof v, out, cap);
  }
  if (n == "lm.state") { int v[5] = {p->lm.cenW, p->lm.cenH, p->lm.cenD, p->lm.frameCount, p->lm.optimized ? 1 : 0}; return put(v, sizeof v, out, cap); }
  if (n == "lm.fullResRegistered") {  // LM.cpp:901-905: laserCloudFullRes through pointAssociateToMap (publish path)
    Cloud reg = p->lo.fullRes;
    lm_register_full_cloud(reg, p->lm.parameters);
    return putv(reg, out, cap);
  }
  if (n == "lm.cornerStack") return putv(p->lm.cornerStack, out, cap);
  if (n == "lm.surfStack") return putv(p->lm.surfStack, out, cap);
  if (n == "lm.cornerFromMap") return putv(p->lm.cornerFromMap, out, cap);
  if (n == "lm.surfFromMap") return putv(p->lm.surfFromMap, out, cap);
  if (n == "lm.validInd") return put(p->lm.validInd, p->lm.validNum * 4, out, cap);
  if (n == "lm.cornerMap") return map_export(p->lm.cornerArray, out, cap);
  if (n == "lm.surfMap") return map_export(p->lm.surfArray, out, cap);
  for (int k = 0; k < 2; ++k) {
    const std::string s = std::to_string(k);
    if (n == "lm.knn.cidx" + s) return putv(p->lm.dbg_cidx[k], out, cap);
    if (n == "lm.knn.sidx" + s) return putv(p->lm.dbg_sidx[k], out, cap);
    if (n == "lm.knn.cd2" + s) return putv(p->lm.dbg_cd2[k], out, cap);
    if (n == "lm.knn.sd2" + s) return putv(p->lm.dbg_sd2[k], out, cap);
    if (n == "lm.knn.cok" + s) return putv(p->lm.dbg_cok[k], out, cap);
    if (n == "lm.knn.sok" + s) return putv(p->lm.dbg_sok[k], out, cap);
  }
  if (n == "lm.costs") {
    double v[4] = {p->lm.dbg_log[0].initial_cost, p->lm.dbg_log[0].final_cost, p->lm.dbg_log[1].initial_cost, p->lm.dbg_log[1].final_cost};
    return put(v, sizeof v, out, cap);
  }
  if (n == "timing") {  // ms: sr lo lm | lm.tree lm.assoc lm.solve lm.filter
    double v[7] = {p->ms_sr, p->ms_lo, p->ms_lm, p->lm.t_tree_ms, p->lm.t_assoc_ms, p->lm.t_solve_ms, p->lm.t_filter_ms};
    return put(v, sizeof v, out, cap);
  }
  return -1;
}

int vloam_oracle_set(void* h, const char* name, const void* data, long bytes) {
  Pipeline* p = (Pipeline*)h;
  const std::string n(name);
  const P4* pts = (const P4*)data; const size_t np = bytes / 16;
  if (n == "lo.last") {  // blob: int nc, int ns, corner points, surf points
    const int* hdr = (const int*)data; const P4* q = (const P4*)((const char*)data + 8);
    p->lo.set_last(Cloud(q, q + hdr[0]), Cloud(q + hdr[0], q + hdr[0] + hdr[1]));
    return 0;
  }
  if (n == "lo.pose") {
    const double* v = (const double*)data;
    memcpy(p->lo.q_w, v, 32); memcpy(p->lo.t_w, v + 4, 24); memcpy(p->lo.para_q, v + 7, 32); memcpy(p->lo.para_t, v + 11, 24);
    return 0;
  }
  if (n == "lm.pose") {
    const double* v = (const double*)data;
    memcpy(p->lm.parameters, v, 56); memcpy(p->lm.q_wmap_wodom, v + 7, 32); memcpy(p->lm.t_wmap_wodom, v + 11, 24);
    return 0;
  }
  if (n == "lm.state") { const int* v = (const int*)data; p->lm.cenW = v[0]; p->lm.cenH = v[1]; p->lm.cenD = v[2]; p->lm.frameCount = v[3]; return 0; }
  if (n == "lm.cornerMap") { map_import(p->lm.cornerArray, data); return 0; }
  if (n == "lm.surfMap") { map_import(p->lm.surfArray, data); return 0; }
  (void)pts; (void)np;
  return -1;
}

// ---- direct access to the restated third-party pieces (unit tests) ---------
int vloam_oracle_voxel_grid(const float* in, int n, float leaf, float* out) {
  Cloud c((const P4*)in, (const P4*)in + n), o;
  voxel_grid(c, leaf, o);
  if (out) memcpy(out, o.data(), o.size() * 16);
  return (int)o.size();
}
int vloam_oracle_knn(const float* cloud, int n, const float* queries, int nq, int k, int backend, int* idx, float* d2) {
  Cloud c((const P4*)cloud, (const P4*)cloud + n);
  KdTree* t = backend ? kd_build(c) : nullptr;
  for (int i = 0; i < nq; ++i) {
    const P4 q = ((const P4*)queries)[i];
    for (int j = 0; j < k; ++j) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = 0; }
    if (t) kd_knn(t, c, q, k, idx + (size_t)i * k, d2 + (size_t)i * k);
    else brute_knn(c, q, k, idx + (size_t)i * k, d2 + (size_t)i * k);
  }
  if (t) kd_free(t);
  return 0;
}
void vloam_oracle_sym_eig3(const double* A, double* evals, double* evecs) { sym_eig3(A, evals, evecs); }
int vloam_oracle_qr_solve_5x3(const double* A, const double* b, double* x) { return colpiv_qr_solve_5x3(A, b, x) ? 1 : 0; }
// n five-point sets (float32[n][5][3]); kind 0: PCA line fit, 1: plane fit.  ok[n], params[n][6] = {a, b} / {normal, d, 0, 0}
void vloam_oracle_fit(const float* near, int n, int kind, int* ok, double* params) {
  for (int i = 0; i < n; ++i) {
    double* o = params + (size_t)i * 6;
    for (int k = 0; k < 6; ++k) o[k] = 0;
    if (kind == 0) ok[i] = fit_line5(near + (size_t)i * 15, o, o + 3) ? 1 : 0;
    else { double d = 0; ok[i] = fit_plane5(near + (size_t)i * 15, o, &d) ? 1 : 0; if (ok[i]) o[3] = d; }
  }
}
// factors: nf x 10 doubles {type, p[3], a[3], b[3]}
static std::vector<Factor> unpack(const double* f, int nf) {
  std::vector<Factor> fs(nf);
  for (int i = 0; i < nf; ++i) {
    fs[i].type = (int)f[i * 10];
    for (int k = 0; k < 3; ++k) { fs[i].p[k] = f[i * 10 + 1 + k]; fs[i].a[k] = f[i * 10 + 4 + k]; fs[i].b[k] = f[i * 10 + 7 + k]; }
  }
  return fs;
}
// the same with a per-factor interpolation ratio s[nf] (DISTORTION == true: slerp(s, q), s * t inside the functors)
static std::vector<Factor> unpack_s(const double* f, const double* s, int nf) {
  std::vector<Factor> fs = unpack(f, nf);
  for (int i = 0; i < nf; ++i) { fs[i].s = s[i]; fs[i].slerp = true; }
  return fs;
}
int vloam_oracle_ceres_solve_s(const double* f, const double* s, int nf, double* x, double* log4) {
  SolveLog lg; ceres_solve(unpack_s(f, s, nf), x, &lg);
  if (log4) { log4[0] = lg.iterations; log4[1] = lg.successful; log4[2] = lg.initial_cost; log4[3] = lg.final_cost; }
  return 0;
}
int vloam_oracle_evaluate_s(const double* f, const double* s, int nf, const double* x, double* cost, double* H, double* g) {
  evaluate_normal_eq(unpack_s(f, s, nf), x, cost, H, g);
  return 0;
}
int vloam_oracle_ceres_solve(const double* f, int nf, double* x, double* log4) {
  SolveLog lg; ceres_solve(unpack(f, nf), x, &lg);
  if (log4) { log4[0] = lg.iterations; log4[1] = lg.successful; log4[2] = lg.initial_cost; log4[3] = lg.final_cost; }
  return 0;
}
int vloam_oracle_evaluate(const double* f, int nf, const double* x, double* cost, double* H, double* g) {
  evaluate_normal_eq(unpack(f, nf), x, cost, H, g);
  return 0;
}
void vloam_oracle_quat(const double* a, const double* b, const double* v, double* ab, double* av, double* ainv) {
  q_mul(a, b, ab); q_rot(a, v, av); q_inv(a, ainv);
}

}  // extern "C"
