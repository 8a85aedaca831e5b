// extern "C" dispatch for the forward / backward entry points (see include/geoconv_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "gcb_common.cuh"
#include "gcb_tc.cuh"

namespace gcb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_yield_sms{0};
int yield_sms() { return g_yield_sms.load(std::memory_order_relaxed); }

// ---- optional kernel timing (one host thread per process drives the library) ----
struct ProfSlot {
  std::vector<cudaEvent_t> start, stop;
  size_t used = 0;
};
static bool g_prof_on = false;
static ProfSlot g_prof[GCB_PROF_N];

void profile_begin(int which, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfSlot& s = g_prof[which];
  if (s.used == s.start.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    s.start.push_back(a);
    s.stop.push_back(b);
  }
  cudaEventRecord(s.start[s.used], st);
}
void profile_end(int which, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfSlot& s = g_prof[which];
  if (s.used < s.start.size()) cudaEventRecord(s.stop[s.used++], st);
}

static int make_rot_table(const int32_t* rot_off_host, int n_rot, int A, RotTable* out) {
  GCB_REQUIRE(rot_off_host, GCB_EINVAL, "rot_off_host is null");
  GCB_REQUIRE(n_rot > 0 && n_rot <= GCB_MAX_ROT, GCB_EINVAL, "n_rot=%d outside [1,%d]", n_rot, GCB_MAX_ROT);
  for (int i = 0; i < GCB_MAX_ROT; ++i) {
    int o = i < n_rot ? rot_off_host[i] : 0;
    out->off[i] = ((o % A) + A) % A;  // torch.roll accepts any integer shift
  }
  return GCB_OK;
}

static int check_shape(int v_out, int v_src, int R, int A, int F, int T) {
  GCB_REQUIRE(v_out >= 0 && v_src > 0 && v_out <= v_src, GCB_EINVAL,
              "need 0 <= v_out (%d) <= v_src (%d)", v_out, v_src);
  GCB_REQUIRE(R > 0 && A > 0 && F > 0 && T > 0, GCB_EINVAL, "R,A,F,T must be positive");
  GCB_REQUIRE((int64_t)v_out * R * A * 3 < (1ll << 31), GCB_EUNSUPPORTED, "V*R*A*3 exceeds int32 slot ids");
  return GCB_OK;
}

}  // namespace gcb

using namespace gcb;

extern "C" int gcb_abi_version(void) { return GCB_ABI_VERSION; }
extern "C" void gcb_yield_sms(int on) { g_yield_sms.store(on ? 1 : 0, std::memory_order_relaxed); }
extern "C" int64_t gcb_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void gcb_profile_enable(int on) {
  g_prof_on = on != 0;
  if (g_prof_on)
    for (int i = 0; i < GCB_PROF_N; ++i) g_prof[i].used = 0;
}
extern "C" int gcb_profile_read(int which, double* ms_total, int64_t* launches) {
  GCB_REQUIRE(which >= 0 && which < GCB_PROF_N && ms_total && launches, GCB_EINVAL, "gcb_profile_read: bad argument");
  ProfSlot& s = g_prof[which];
  double total = 0.0;
  for (size_t i = 0; i < s.used; ++i) {
    GCB_CUDA_OK(cudaEventSynchronize(s.stop[i]));
    float ms = 0.f;
    GCB_CUDA_OK(cudaEventElapsedTime(&ms, s.start[i], s.stop[i]));
    total += ms;
  }
  *ms_total = total;
  *launches = (int64_t)s.used;
  return GCB_OK;
}
extern "C" const char* gcb_last_error(void) { return g_err; }

extern "C" int gcb_tc_supported(int32_t R, int32_t A, int32_t F, int32_t T, int32_t n_rot) {
  return tc_supported(R, A, F, T, n_rot) ? 1 : 0;
}

extern "C" int gcb_tc_half_supported(int32_t R, int32_t A, int32_t F, int32_t T, int32_t n_rot) {
  return tc_supported(R, A, F, T, n_rot) && tc_half_supported(R, A, F, T, n_rot) ? 1 : 0;
}

extern "C" size_t gcb_conv_fwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A,
                                               int32_t F, int32_t T, int32_t n_rot, int precision) {
  (void)v_src;
  size_t y_tmp = align_up((size_t)v_out * n_rot * T * sizeof(float), 256);  // when caller passes y == NULL
  if (precision == GCB_PREC_FP32) return y_tmp + ffma_fwd_workspace_bytes(v_out, R, A, F) + 256;
  size_t tc = tc_fwd_workspace_bytes(v_out, R, A, F, T, n_rot, precision);
  if (tc_centre_fwd_supported(F, T) && tc_centre_fwd_bytes(v_out, F, T) > tc) tc = tc_centre_fwd_bytes(v_out, F, T);
  return y_tmp + tc + 256;   // (a centre-only layer on the CUDA-core path uses no workspace)
}

extern "C" int gcb_conv_fwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                            const float* w_self, const float* bias, const int32_t* rot_off_host,
                            int32_t n_rot, int32_t v_out, int32_t v_src, int32_t R, int32_t A,
                            int32_t F, int32_t T, int act, int precision, float* y, float* z,
                            int32_t* argmax, float* interp_out, const void* records, const void* tables,
                            uint32_t* stats_out, int32_t row0, void* workspace, size_t workspace_bytes,
                            void* stream) {
  GCB_REQUIRE(x && idx && w && w_self && bias, GCB_EINVAL, "gcb_conv_fwd: null input pointer");
  GCB_REQUIRE(y || z, GCB_EINVAL, "gcb_conv_fwd: need y and/or z");
  GCB_REQUIRE((z == nullptr) == (argmax == nullptr), GCB_EINVAL, "gcb_conv_fwd: z and argmax go together");
  GCB_REQUIRE(act >= 0 && act <= GCB_ACT_TANH, GCB_EINVAL, "gcb_conv_fwd: unknown activation %d", act);
  GCB_REQUIRE(row0 >= 0 && (int64_t)row0 + v_out <= v_src, GCB_EINVAL,
              "gcb_conv_fwd: rows [row0, row0 + v_out) = [%d, %lld) must lie inside the signal (%d rows)", row0,
              (long long)row0 + v_out, v_src);
  int rc = check_shape(v_out, v_src, R, A, F, T);
  if (rc) return rc;
  RotTable rot;
  rc = make_rot_table(rot_off_host, n_rot, A, &rot);
  if (rc) return rc;
  if (v_out == 0) return GCB_OK;
  GCB_REQUIRE(workspace && workspace_bytes >= gcb_conv_fwd_workspace_bytes(v_out, v_src, R, A, F, T, n_rot, precision),
              GCB_EWORKSPACE, "gcb_conv_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  size_t y_tmp = align_up((size_t)v_out * n_rot * T * sizeof(float), 256);
  float* y_out = y ? y : (float*)ws;
  ws += y_tmp;
  workspace_bytes -= y_tmp;
  if (stats_out && (precision != GCB_PREC_F16X3 || !w_eff)) GCB_CUDA_OK(cudaMemsetAsync(stats_out, 0, 32, st));
  if (!w_eff && precision != GCB_PREC_FP32 && tc_centre_fwd_supported(F, T) && ((uintptr_t)x % 16 == 0)) {
    // centre-only (all-zero prior): one tensor-core GEMM, bias / activation / pooling in its epilogue kernel
    return tc_centre_fwd(x + (size_t)row0 * F, w_self, bias, n_rot, v_out, F, T, act, precision, y, z, argmax, ws,
                         workspace_bytes, st);
  }
  if (precision == GCB_PREC_FP32 || !w_eff) {
    rc = ffma_conv_fwd(x, idx, w, w_eff, w_self, bias, rot, n_rot, v_out, v_src, R, A, F, T, act, y_out, ws, st, row0);
  } else if (precision == GCB_PREC_TF32 || precision == GCB_PREC_TF32X3 || precision == GCB_PREC_F16X3) {
    GCB_REQUIRE(tc_supported(R, A, F, T, n_rot), GCB_EUNSUPPORTED,
                "gcb_conv_fwd: tcgen05 path does not support R=%d A=%d F=%d T=%d n_rot=%d", R, A, F, T, n_rot);
    GCB_REQUIRE(precision != GCB_PREC_F16X3 || tc_half_supported(R, A, F, T, n_rot), GCB_EUNSUPPORTED,
                "gcb_conv_fwd: the fp16 split does not support R=%d A=%d F=%d T=%d n_rot=%d", R, A, F, T, n_rot);
    // the tensor-core path pools in its own finishing kernel
    return tc_conv_fwd(x, idx, w, w_eff, w_self, bias, rot, n_rot, v_out, v_src, R, A, F, T, act, precision,
                       y_out, z, argmax, interp_out, records, tables, stats_out, row0, ws, workspace_bytes, st);
  } else {
    GCB_REQUIRE(false, GCB_EINVAL, "gcb_conv_fwd: unknown precision %d", precision);
  }
  if (rc) return rc;
  if (z) rc = amp_fwd(y_out, v_out, n_rot, T, z, argmax, st);
  return rc;
}

extern "C" size_t gcb_fwd_records_bytes(int32_t v_out, int32_t R, int32_t A) { return tc_fwd_records_bytes(v_out, R, A); }

extern "C" int gcb_fwd_pack_records(const int32_t* idx, const float* w, int32_t v_out, int32_t R, int32_t A,
                                    void* records, void* stream) {
  GCB_REQUIRE(idx && w && records, GCB_EINVAL, "gcb_fwd_pack_records: null pointer");
  GCB_REQUIRE(v_out > 0 && R > 0 && A > 0, GCB_EINVAL, "gcb_fwd_pack_records: bad sizes");
  return tc_fwd_pack_records(idx, w, v_out, R, A, records, (cudaStream_t)stream);
}

extern "C" size_t gcb_fwd_tables_bytes(void) { return tc_fwd_tables_bytes(); }

extern "C" int gcb_fwd_build_tables(const int32_t* rot_off_host, int32_t n_rot, int32_t v_out, int32_t R, int32_t A,
                                    int32_t F, int32_t T, int precision, void* tables, void* stream) {
  GCB_REQUIRE(tables, GCB_EINVAL, "gcb_fwd_build_tables: null pointer");
  GCB_REQUIRE(precision != GCB_PREC_FP32 && tc_supported(R, A, F, T, n_rot), GCB_EUNSUPPORTED,
              "gcb_fwd_build_tables: only the tcgen05 forward uses tables");
  RotTable rot;
  int rc = make_rot_table(rot_off_host, n_rot, A, &rot);
  if (rc) return rc;
  return tc_fwd_build_tables(rot, n_rot, v_out, R, A, F, T, precision, tables, (cudaStream_t)stream);
}

extern "C" size_t gcb_conv_bwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A,
                                               int32_t F, int32_t T, int precision) {
  size_t d_full = align_up((size_t)v_out * R * A * F * sizeof(float), 256);
  size_t ffma = d_full + ffma_bwd_workspace_bytes(v_out, R, A, F) + 256;
  if (precision != GCB_PREC_FP32 && tc_bwd_supported(R, A, F, T)) {
    size_t tc = tc_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision) + 256;
    return tc > ffma ? tc : ffma;
  }
  return ffma;
}

extern "C" int gcb_conv_bwd_workspace_selfcheck(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F,
                                                int32_t T, int precision) {
  if (precision == GCB_PREC_FP32 || !tc_bwd_supported(R, A, F, T)) return 1;   // CUDA-core route: one fixed layout
  return tc_bwd_layout_fits(v_out, v_src, R, A, F, T, precision) &&
                 gcb_conv_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision) >=
                     tc_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision)
             ? 1 : 0;
}

extern "C" int gcb_conv_bwd(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                            const float* w_self, const float* h, const int32_t* rot_sel,
                            const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags, int32_t v_out,
                            int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T, int precision,
                            int accumulate, const float* interp, float* d_x, float* d_w_eff, float* d_w_self,
                            float* d_bias, void* workspace, size_t workspace_bytes, void* stream) {
  GCB_REQUIRE(x && idx && w && w_self && h && rot_sel, GCB_EINVAL, "gcb_conv_bwd: null input pointer");
  GCB_REQUIRE(!d_x || !w_eff || (csr_row_ptr && csr_entries), GCB_EINVAL, "gcb_conv_bwd: d_x needs the CSR");
  GCB_REQUIRE(d_w_eff && d_w_self && d_bias, GCB_EINVAL, "gcb_conv_bwd: null output pointer");
  int rc = check_shape(v_out, v_src, R, A, F, T);
  if (rc) return rc;
  GCB_REQUIRE(v_out > 0, GCB_EINVAL, "gcb_conv_bwd: empty input");
  GCB_REQUIRE(workspace && workspace_bytes >= gcb_conv_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision),
              GCB_EWORKSPACE, "gcb_conv_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* d_full = (float*)ws;
  ws += align_up((size_t)v_out * R * A * F * sizeof(float), 256);
  if (precision != GCB_PREC_FP32 && tc_bwd_supported(R, A, F, T))   // (w_eff == NULL: centre-only GEMMs, no CSR needed)
    return tc_conv_bwd(x, idx, w, w_eff, w_self, h, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, F, T,
                       precision, accumulate, interp, d_x, d_w_eff, d_w_self, d_bias, workspace, workspace_bytes, st,
                       nullptr, (csr_flags & GCB_CSR_RING_GROUPED) != 0);
  rc = ffma_conv_bwd(x, idx, w, w_eff, h, rot_sel, v_out, v_src, R, A, F, T, accumulate, d_w_eff, d_w_self,
                     d_bias, d_x ? d_full : nullptr, ws, st);
  if (rc || !d_x) return rc;
  return csr_reduce_dx(w_eff ? d_full : nullptr, w, h, w_self, rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, F, T,
                       accumulate, d_x, st);
}

// ---- dense backward: a gradient for EVERY rotation (the layer used without AngularMaxPooling, or under min /
// average pooling) in ONE pass through G instead of n_rot sparse backward calls --------------------------------
extern "C" int gcb_conv_bwd_dense_supported(int32_t R, int32_t A, int32_t F, int32_t T, int precision) {
  return precision != GCB_PREC_FP32 && tc_bwd_supported(R, A, F, T) && A <= GCB_MAX_ROT && T % 32 == 0 ? 1 : 0;
}

extern "C" int gcb_conv_bwd_dense(const float* x, const int32_t* idx, const float* w, const float* w_eff,
                                  const float* w_self, const float* h_all, const int32_t* rot_off_host, int32_t n_rot,
                                  const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags, int32_t v_out,
                                  int32_t v_src, int32_t R, int32_t A, int32_t F, int32_t T, int precision,
                                  int accumulate, float* d_x, float* d_w_eff, float* d_w_self, float* d_bias,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  GCB_REQUIRE(x && idx && w && w_eff && w_self && h_all && rot_off_host && csr_row_ptr && csr_entries, GCB_EINVAL,
              "gcb_conv_bwd_dense: null input pointer");
  GCB_REQUIRE(d_w_eff && d_w_self && d_bias, GCB_EINVAL, "gcb_conv_bwd_dense: null output pointer");
  int rc = check_shape(v_out, v_src, R, A, F, T);
  if (rc) return rc;
  GCB_REQUIRE(v_out > 0 && n_rot > 0 && n_rot <= GCB_MAX_ROT, GCB_EINVAL, "gcb_conv_bwd_dense: bad sizes");
  GCB_REQUIRE(gcb_conv_bwd_dense_supported(R, A, F, T, precision) && (csr_flags & GCB_CSR_RING_GROUPED), GCB_EUNSUPPORTED,
              "gcb_conv_bwd_dense: needs the tensor-core backward, a ring-grouped CSR, A <= %d and T %% 32 == 0", GCB_MAX_ROT);
  GCB_REQUIRE(workspace && workspace_bytes >= gcb_conv_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision),
              GCB_EWORKSPACE, "gcb_conv_bwd_dense: workspace too small");
  return tc_conv_bwd(x, idx, w, w_eff, w_self, h_all, nullptr, csr_row_ptr, csr_entries, v_out, v_src, R, A, F, T, precision,
                     accumulate, nullptr, d_x, d_w_eff, d_w_self, d_bias, workspace, workspace_bytes, (cudaStream_t)stream,
                     nullptr, true, 3, n_rot, rot_off_host);
}

// ---- AMP-sparse backward from (dZ, Z, argmax): one call, one preparation kernel ---------------------------------
extern "C" size_t gcb_conv_amp_bwd_workspace_bytes(int32_t v_out, int32_t v_src, int32_t R, int32_t A, int32_t F,
                                                   int32_t T, int precision) {
  return gcb_conv_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision) +
         align_up((size_t)v_out * T * sizeof(float), 256) + align_up((size_t)v_out * sizeof(int32_t), 256) + 512;
}

extern "C" int gcb_conv_amp_bwd_can_split(int32_t R, int32_t A, int32_t F, int32_t T, int precision, int has_neighbor) {
  return precision != GCB_PREC_FP32 && tc_bwd_supported(R, A, F, T) && has_neighbor && T % 4 == 0 && T <= 1024 ? 1 : 0;
}

namespace gcb {
int act_grad_launch(const float* grad, const float* out, int64_t n, int act, float* h, cudaStream_t st);
int select_rot_launch(const int32_t* argmax, const RotTable& rot, int n_rot, int v, int32_t* rot_sel, cudaStream_t st);
}  // namespace gcb

extern "C" int gcb_conv_amp_bwd(const float* grad_z, const float* z, const int32_t* argmax,
                                const int32_t* rot_off_host, int32_t n_rot, int act, const float* x,
                                const int32_t* idx, const float* w, const float* w_eff, const float* w_self,
                                const int32_t* csr_row_ptr, const int32_t* csr_entries, int csr_flags,
                                const uint32_t* csr_stats, const uint32_t* fwd_stats, int32_t v_out, int32_t v_src, int32_t R, int32_t A,
                                int32_t F, int32_t T, int precision, int stages, float* d_x, float* d_w_eff,
                                float* d_w_self, float* d_bias, void* workspace, size_t workspace_bytes,
                                void* stream) {
  GCB_REQUIRE(stages >= 1 && stages <= 3, GCB_EINVAL, "gcb_conv_amp_bwd: stages must be 1, 2 or 3");
  GCB_REQUIRE(stages == 3 || gcb_conv_amp_bwd_can_split(R, A, F, T, precision, w_eff != nullptr), GCB_EUNSUPPORTED,
              "gcb_conv_amp_bwd: this shape / precision runs in one stage only");
  GCB_REQUIRE(grad_z && z && argmax && x && idx && w && w_self, GCB_EINVAL, "gcb_conv_amp_bwd: null input pointer");
  GCB_REQUIRE(!d_x || !w_eff || (csr_row_ptr && csr_entries), GCB_EINVAL, "gcb_conv_amp_bwd: d_x needs the CSR");
  GCB_REQUIRE(d_w_eff && d_w_self && d_bias, GCB_EINVAL, "gcb_conv_amp_bwd: null output pointer");
  GCB_REQUIRE(act >= 0 && act <= GCB_ACT_TANH, GCB_EINVAL, "gcb_conv_amp_bwd: unknown activation %d", act);
  int rc = check_shape(v_out, v_src, R, A, F, T);
  if (rc) return rc;
  GCB_REQUIRE(v_out > 0, GCB_EINVAL, "gcb_conv_amp_bwd: empty input");
  GCB_REQUIRE(workspace && workspace_bytes >= gcb_conv_amp_bwd_workspace_bytes(v_out, v_src, R, A, F, T, precision),
              GCB_EWORKSPACE, "gcb_conv_amp_bwd: workspace too small");
  TcBwdAmp amp;
  rc = make_rot_table(rot_off_host, n_rot, A, &amp.rot);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)align_up((size_t)workspace, 256);
  amp.h = (float*)ws;
  ws += align_up((size_t)v_out * T * sizeof(float), 256);
  amp.rot_sel = (int32_t*)ws;
  ws += align_up((size_t)v_out * sizeof(int32_t), 256);
  const size_t rest = workspace_bytes - (size_t)(ws - (char*)workspace);
  amp.grad_z = grad_z; amp.z = z; amp.argmax = argmax; amp.n_rot = n_rot; amp.act = act;
  amp.csr_stats = csr_stats; amp.fwd_stats = fwd_stats;
  const bool aligned = ((uintptr_t)grad_z % 16 == 0) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)d_bias % 16 == 0) && T % 4 == 0 && T <= 1024;
  if (precision != GCB_PREC_FP32 && tc_bwd_supported(R, A, F, T) && aligned)
    return tc_conv_bwd(x, idx, w, w_eff, w_self, amp.h, amp.rot_sel, csr_row_ptr, csr_entries, v_out, v_src, R, A, F, T,
                       precision, 0, nullptr, d_x, d_w_eff, d_w_self, d_bias, ws, rest, st, &amp,
                       (csr_flags & GCB_CSR_RING_GROUPED) != 0, stages);
  // CUDA-core route (or a shape the preparation kernel does not tile): the separate small kernels, then gcb_conv_bwd
  rc = act_grad_launch(grad_z, z, (int64_t)v_out * T, act, amp.h, st);
  if (rc) return rc;
  rc = select_rot_launch(argmax, amp.rot, n_rot, v_out, amp.rot_sel, st);
  if (rc) return rc;
  return gcb_conv_bwd(x, idx, w, w_eff, w_self, amp.h, amp.rot_sel, csr_row_ptr, csr_entries, csr_flags, v_out, v_src, R,
                      A, F, T, precision, 0, nullptr, d_x, d_w_eff, d_w_self, d_bias, ws, rest, stream);
}
