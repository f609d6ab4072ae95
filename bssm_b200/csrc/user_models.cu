// user_models.cu -- model functors supplied by the caller as CUDA C++ source, compiled at run time.
//
// The reference defines ssm_sde (and ssm_nlg) models through C++ function pointers that the user
// compiles with Rcpp::sourceCpp and hands over as externalptr (src/model_ssm_sde.h:8-11:
//   double drift(x, theta), diffusion(x, theta), ddiffusion(x, theta), log_obs_density(y, alpha, theta);
// vignettes/sde_model.Rmd).  A host function pointer cannot be called from a kernel, so the same four
// functions are taken as SOURCE TEXT with a fixed scalar signature
//   __device__ double drift(double x, const double* theta);
//   __device__ double diffusion(double x, const double* theta);
//   __device__ double ddiffusion(double x, const double* theta);
//   __device__ double log_obs_density(double y, double x, const double* theta);
// and compiled by NVRTC into the SAME filter kernels the registry models use (pf::filter_kernel /
// pf::small_filter_kernel over pf::SdeBsf<UserSde>): this translation unit writes the adapter, asks
// NVRTC for the kernel variant a launch needs (lazily, one variant per compilation, cached), loads the
// cubin with cudaLibraryLoadData and launches the cudaKernel_t like any other kernel.  The user code
// can call the deterministic math of include/bssm_b200/pf_math.h (pf_exp, pf_log, pf_log1p, ...).
//
// ssm_nlg (bootstrap filter, p = 1; src/model_ssm_nlg.h:10-21: Z_fn, H_fn, T_fn, R_fn, a1_fn, P1_fn with
// arma::vec / arma::mat returns) the same way, with the outputs written through a pointer:
//   __device__ double Z_fn(int t, const double* alpha, const double* theta, const double* known_params);
//   __device__ double H_fn(int t, const double* alpha, const double* theta, const double* known_params);   // sd
//   __device__ void   T_fn(int t, const double* alpha, const double* theta, const double* known_params, double* out); // m
//   __device__ void   R_fn(int t, const double* alpha, const double* theta, const double* known_params, double* out); // m x k
//   __device__ void   a1_fn(const double* theta, const double* known_params, double* out);                // m
//   __device__ void   P1_fn(const double* theta, const double* known_params, double* out);                // m x m
// (matrices column-major; the reference's known_tv_params go to the tail of known_params, indexed with t).
// The Jacobians Z_gn / T_gn are not needed by the bootstrap filter; the psi-APF and the EKF particle
// filter of user models need the host-side approximation and stay with the registry models.
// libnvrtc is dlopen'ed on first use: the library does not depend on it otherwise.
#include "filters.h"
#include <dlfcn.h>
#include <nvrtc.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace pfl {
namespace {

struct NvrtcApi {
  void* h = nullptr;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
  bool load(std::string& why) {
    if (h) return true;
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names)
      if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) { why = "libnvrtc not found (run-time compilation of user models needs it)"; return false; }
#define PFL_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(h, "nvrtc" #f)); if (!f) { why = "libnvrtc: missing symbol nvrtc" #f; h = nullptr; return false; }
    PFL_SYM(CreateProgram) PFL_SYM(DestroyProgram) PFL_SYM(AddNameExpression) PFL_SYM(CompileProgram)
    PFL_SYM(GetProgramLogSize) PFL_SYM(GetProgramLog) PFL_SYM(GetLoweredName) PFL_SYM(GetCUBINSize)
    PFL_SYM(GetCUBIN) PFL_SYM(GetErrorString)
#undef PFL_SYM
    return true;
  }
};

enum { USER_SDE = 0, USER_NLG = 1 };
struct UserModel {
  int kind = USER_SDE;
  std::string source;
  std::map<std::string, cudaKernel_t> kernels; // kernel name expression -> loaded kernel (this device context)
  std::vector<cudaLibrary_t> libs;
};

std::mutex g_mu;
NvrtcApi g_rtc;
std::vector<std::unique_ptr<UserModel>> g_models;
thread_local std::string g_uerr;

// directory of this shared library: the kernel headers NVRTC includes live next to it (in-tree build)
std::string lib_dir() {
  Dl_info info;
  if (!dladdr(reinterpret_cast<void*>(&lib_dir), &info) || !info.dli_fname) return ".";
  std::string p = info.dli_fname;
  const size_t s = p.find_last_of('/');
  return s == std::string::npos ? "." : p.substr(0, s);
}

std::string sde_program(const std::string& user) {
  return std::string(
             "#include \"models_nlg_sde.cuh\"\n#include \"pf_small.cuh\"\n"
             "namespace bssm_user {\n#line 1 \"user_model\"\n") +
         user +
         "\n}\nnamespace pf {\n"
         "struct UserSde {\n"
         "  static __device__ __forceinline__ double drift(const SdeParams& mp, double x) { return bssm_user::drift(x, mp.theta_u); }\n"
         "  static __device__ __forceinline__ double diffusion(const SdeParams& mp, double x) { return bssm_user::diffusion(x, mp.theta_u); }\n"
         "  static __device__ __forceinline__ double ddiffusion(const SdeParams& mp, double x) { return bssm_user::ddiffusion(x, mp.theta_u); }\n"
         "  static __device__ __forceinline__ double log_obs(const SdeParams& mp, double y, double, double x) { return bssm_user::log_obs_density(y, x, mp.theta_u); }\n"
         "};\n}\n";
}

std::string nlg_program(const std::string& user) {
  return std::string(
             "#include \"models_nlg_sde.cuh\"\n#include \"pf_small.cuh\"\n"
             "namespace bssm_user {\n#line 1 \"user_model\"\n") +
         user +
         "\n}\nnamespace pf {\n"
         "struct UserNlg {\n"
         "  static __device__ __forceinline__ double Z(int t, const double* a, const double* th, const double* kp) { return bssm_user::Z_fn(t, a, th, kp); }\n"
         "  static __device__ __forceinline__ double H(int t, const double* a, const double* th, const double* kp) { return bssm_user::H_fn(t, a, th, kp); }\n"
         "  static __device__ __forceinline__ void T(int t, const double* a, const double* th, const double* kp, double* o) { bssm_user::T_fn(t, a, th, kp, o); }\n"
         "  static __device__ __forceinline__ void R(int t, const double* a, const double* th, const double* kp, double* o) { bssm_user::R_fn(t, a, th, kp, o); }\n"
         "  static __device__ __forceinline__ void a1(const double* th, const double* kp, double* o) { bssm_user::a1_fn(th, kp, o); }\n"
         "  static __device__ __forceinline__ void P1(const double* th, const double* kp, double* o) { bssm_user::P1_fn(th, kp, o); }\n"
         "};\n}\n";
}

// compiles `expr` (a __global__ template instantiation) of the model's program and loads it
int get_kernel(UserModel& um, const std::string& expr, cudaKernel_t* out) {
  auto it = um.kernels.find(expr);
  if (it != um.kernels.end()) { *out = it->second; return 0; }
  if (!g_rtc.load(g_uerr)) return -2;
  const std::string prog_src = um.kind == USER_NLG ? nlg_program(um.source) : sde_program(um.source);
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_rtc.CreateProgram(&prog, prog_src.c_str(), "bssm_b200_user_model.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) { g_uerr = std::string("nvrtcCreateProgram: ") + g_rtc.GetErrorString(r); return -2; }
  struct PG { nvrtcProgram* p; ~PG() { g_rtc.DestroyProgram(p); } } pg{&prog};
  g_rtc.AddNameExpression(prog, expr.c_str());
  const std::string dir = lib_dir();
  const std::string i1 = "-I" + dir, i2 = "-I" + dir + "/../../include";
  // the flags of bssm_b200/build.py that matter for the bits: no FMA contraction
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", i1.c_str(), i2.c_str()};
  r = g_rtc.CompileProgram(prog, 6, opts);
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    g_rtc.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) g_rtc.GetProgramLog(prog, &log[0]);
    g_uerr = std::string("user model does not compile (") + g_rtc.GetErrorString(r) + "):\n" + log;
    return -2;
  }
  const char* lowered = nullptr;
  if ((r = g_rtc.GetLoweredName(prog, expr.c_str(), &lowered)) != NVRTC_SUCCESS || !lowered) {
    g_uerr = "nvrtcGetLoweredName failed for " + expr;
    return -2;
  }
  size_t sz = 0;
  g_rtc.GetCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  if (!sz || g_rtc.GetCUBIN(prog, cubin.data()) != NVRTC_SUCCESS) { g_uerr = "nvrtcGetCUBIN failed"; return -2; }
  cudaLibrary_t lib = nullptr;
  PFL_CU(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t k = nullptr;
  cudaError_t e = cudaLibraryGetKernel(&k, lib, lowered);
  if (e != cudaSuccess) { cudaLibraryUnload(lib); return (int)e; }
  um.libs.push_back(lib);
  um.kernels[expr] = k;
  *out = k;
  return 0;
}

std::string filter_kernel_expr(const std::string& model, const Geometry& g);
std::string sde_kernel_expr(const Geometry& g, bool per_filter) {
  return filter_kernel_expr(std::string("pf::SdeBsf<pf::UserSde, ") + (per_filter ? "true" : "false") + ">", g);
}
std::string nlg_model_expr(int m, int k, bool per_filter) {
  return "pf::NlgFnBsf<pf::UserNlg, " + std::to_string(m) + ", " + std::to_string(k) + ", " + (per_filter ? "true" : "false") + ">";
}
std::string filter_kernel_expr(const std::string& model, const Geometry& g) {
  if (g.sync_mode == pf::SYNC_SMALL)
    return "pf::small_filter_kernel<" + model + ", " + std::to_string(g.P) + ", " + (g.block <= 512 ? "512" : "1024") + ">";
  const bool small = g.block <= 512;
  return "pf::filter_kernel<" + model + ", " + std::to_string(g.P) + ", " +
         (small ? std::to_string(PF_SMALL_MAXT) + ", " + std::to_string(PF_SMALL_MINB) : std::string("1024, 1")) + ">";
}

UserModel* find(int model_id) {
  const int i = model_id - 1000;
  if (i < 0 || i >= (int)g_models.size()) return nullptr;
  return g_models[i].get();
}

} // namespace

const char* user_model_error() { return g_uerr.c_str(); }
// bssm_b200_shutdown: unload the compiled kernels; the registered sources (and their ids) stay valid and
// are compiled again on the next use
void user_models_release() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& um : g_models) {
    for (cudaLibrary_t lib : um->libs) cudaLibraryUnload(lib);
    um->libs.clear();
    um->kernels.clear();
  }
}
bool user_sde_exists(int model_id) {
  std::lock_guard<std::mutex> lk(g_mu);
  UserModel* um = find(model_id);
  return um && um->kind == USER_SDE;
}

// registers the source and compiles the latency-kernel variant at once, so that errors in the user's
// code surface here and not at the first filter call
int user_sde_register(const char* source, int* model_id) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_uerr.clear();
  for (size_t i = 0; i < g_models.size(); ++i)
    if (g_models[i]->kind == USER_SDE && g_models[i]->source == source) { *model_id = 1000 + (int)i; return 0; }
  std::unique_ptr<UserModel> um(new UserModel());
  um->source = source;
  Geometry g;
  g.P = 2; g.block = 64; g.ranks = 1; g.sync_mode = pf::SYNC_SMALL;
  cudaKernel_t k;
  int rc = get_kernel(*um, sde_kernel_expr(g, false), &k);
  if (rc) return rc;
  g_models.push_back(std::move(um));
  *model_id = 1000 + (int)g_models.size() - 1;
  return 0;
}

static int launch_kernel(cudaKernel_t k, void** args, const Geometry& g, uint32_t nf, size_t small_smem, Stream stream);
int launch_user_sde(int model_id, const pf::FilterParams& fp, const pf::SdeParams& mp, const Geometry& g, uint32_t nf,
                    bool per_filter, Stream stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_uerr.clear();
  UserModel* um = find(model_id);
  if (!um || um->kind != USER_SDE) { g_uerr = "unknown ssm_sde user model id"; return -2; }
  cudaKernel_t k;
  int rc = get_kernel(*um, sde_kernel_expr(g, per_filter), &k);
  if (rc) return rc;
  void* args[2] = {const_cast<pf::FilterParams*>(&fp), const_cast<pf::SdeParams*>(&mp)};
  // same state dimension and noise record as the registry policies
  const size_t smem = pf::small_smem_bytes<pf::SdeGbmBsf, 2>(fp.N);
  return launch_kernel(k, args, g, nf, smem, stream);
}

// shared-memory bytes of the latency kernel for an (m, k) Gaussian-step policy (pf_small.cuh)
static size_t small_smem_mk(uint32_t N, int m, int k, int P) {
  const size_t CT = pf::small_consumers(N, P), Np = (size_t)P * CT;
  return 2 * (size_t)m * Np * 8 + 2 * Np * 8 + 2 * CT * ((size_t)P * k * 8) + Np * 4;
}

int user_nlg_register(const char* source, int m, int k, int* model_id) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_uerr.clear();
  for (size_t i = 0; i < g_models.size(); ++i)
    if (g_models[i]->kind == USER_NLG && g_models[i]->source == source) { *model_id = 1000 + (int)i; return 0; }
  std::unique_ptr<UserModel> um(new UserModel());
  um->kind = USER_NLG;
  um->source = source;
  Geometry g;
  g.P = 2; g.block = 64; g.ranks = 1; g.sync_mode = pf::SYNC_SMALL;
  cudaKernel_t kk;
  int rc = get_kernel(*um, filter_kernel_expr(nlg_model_expr(m, k, false), g), &kk); // compile errors surface here
  if (rc) return rc;
  g_models.push_back(std::move(um));
  *model_id = 1000 + (int)g_models.size() - 1;
  return 0;
}
bool user_nlg_exists(int model_id) {
  std::lock_guard<std::mutex> lk(g_mu);
  UserModel* um = find(model_id);
  return um && um->kind == USER_NLG;
}

int launch_user_nlg(int model_id, int m, int k, const pf::FilterParams& fp, const pf::NlgParams& mp, const Geometry& g,
                    uint32_t nf, bool per_filter, Stream stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_uerr.clear();
  UserModel* um = find(model_id);
  if (!um || um->kind != USER_NLG) { g_uerr = "unknown ssm_nlg user model id"; return -2; }
  cudaKernel_t kk;
  int rc = get_kernel(*um, filter_kernel_expr(nlg_model_expr(m, k, per_filter), g), &kk);
  if (rc) return rc;
  void* args[2] = {const_cast<pf::FilterParams*>(&fp), const_cast<pf::NlgParams*>(&mp)};
  return launch_kernel(kk, args, g, nf, small_smem_mk(fp.N, m, k, g.P), stream);
}

// a1_fn / P1_fn on the device for nf parameter vectors: out (device) = nf x (m + m m)
int user_nlg_init(int model_id, int m, const double* theta, size_t stride, const double* known, uint32_t nf, double* out,
                  Stream stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_uerr.clear();
  UserModel* um = find(model_id);
  if (!um || um->kind != USER_NLG) { g_uerr = "unknown ssm_nlg user model id"; return -2; }
  cudaKernel_t kk;
  int rc = get_kernel(*um, "pf::nlg_user_init_kernel<pf::UserNlg, " + std::to_string(m) + ">", &kk);
  if (rc) return rc;
  void* args[5] = {&theta, &stride, &known, &nf, &out};
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((nf + 127) / 128, 1, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.stream = stream;
  PFL_CU(cudaLaunchKernelExC(&cfg, (const void*)kk, args));
  return 0;
}

int occupancy_user_nlg(int model_id, int m, int k, const Geometry& g, bool per_filter, int* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  UserModel* um = find(model_id);
  if (!um || um->kind != USER_NLG) { g_uerr = "unknown ssm_nlg user model id"; return -2; }
  cudaKernel_t kk;
  int rc = get_kernel(*um, filter_kernel_expr(nlg_model_expr(m, k, per_filter), g), &kk);
  if (rc) return rc;
  int dev = 0, sms = 0, per_sm = 0;
  PFL_CU(cudaGetDevice(&dev));
  PFL_CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PFL_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kk, g.block, 0));
  *out = g.sync_mode == pf::SYNC_GROUP ? (per_sm * sms) / (int)g.ranks : per_sm * sms;
  return 0;
}

static int launch_kernel(cudaKernel_t k, void** args, const Geometry& g, uint32_t nf, size_t small_smem, Stream stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nf * g.ranks, 1, 1);
  cfg.blockDim = dim3(g.block, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (g.sync_mode == pf::SYNC_SMALL) {
    cfg.dynamicSmemBytes = small_smem;
    if (small_smem > 48 * 1024) PFL_CU(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem));
  } else if (g.sync_mode == pf::SYNC_CLUSTER) {
    if (g.ranks > 8) PFL_CU(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.ranks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
  } else if (g.sync_mode == pf::SYNC_GROUP) {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  }
  PFL_CU(cudaLaunchKernelExC(&cfg, (const void*)k, args));
  return 0;
}

// co-resident filters of this geometry (SYNC_GROUP waves)
int occupancy_user_sde(int model_id, const Geometry& g, bool per_filter, int* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  UserModel* um = find(model_id);
  if (!um || um->kind != USER_SDE) { g_uerr = "unknown ssm_sde user model id"; return -2; }
  cudaKernel_t k;
  int rc = get_kernel(*um, sde_kernel_expr(g, per_filter), &k);
  if (rc) return rc;
  int dev = 0, sms = 0, per_sm = 0;
  PFL_CU(cudaGetDevice(&dev));
  PFL_CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PFL_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, g.block, 0));
  *out = g.sync_mode == pf::SYNC_GROUP ? (per_sm * sms) / (int)g.ranks : per_sm * sms;
  return 0;
}

} // namespace pfl
