// Device side of the engine as seen by the C ABI (capi.cu).  Implemented in chol_gpu.cu / klu_gpu.cu.
#pragma once
#include "host.hpp"
#include <string>
#include <vector>

namespace b200s {

// status codes shared with include/b200sparse.h (kept numerically identical)
enum : int { ST_OK = 0, ST_NOT_POSDEF = 1, ST_SINGULAR = 2, ST_OOM = -2, ST_TOO_LARGE = -3, ST_INVALID = -4,
             ST_NO_DEVICE = -5, ST_CUDA = -6 };

void set_last_error(const std::string& s);
const char* get_last_error();
int  current_device();          // device chosen with b200s_set_device (default 0)
void set_current_device(int d);
int  device_count();

struct CholTimes {
    double ms_h2d = 0, ms_assemble = 0, ms_factor = 0, ms_total = 0, ms_solve = 0;
    double ms_dense_update = 0, ms_potrf = 0, ms_trsm = 0, ms_extend = 0;
};

class CholDevice;   // device-resident plan + numeric factor
CholDevice* chol_device_create(const CholPlan& plan, const CholOpts& opts, int device, int* status);
void chol_device_destroy(CholDevice* d);
int  chol_device_factorize(CholDevice* d, const double* val, bool val_on_device, i64* minor, CholTimes* times);
int  chol_device_solve(CholDevice* d, int sys, double* B, i64 nrhs, i64 ldB, bool on_device, CholTimes* times);
// device-resident right-hand sides on the handle's stream, no synchronisation; sys 9 = L x = P b, 10 = x = P' L^-T b
int  chol_device_solve_async(CholDevice* d, int sys, double* B_dev, i64 nrhs, i64 ldB);
// sparse right-hand sides (n x ncols CCS, sorted rows): structure-aware solve -- sparse upload, forward sweep restricted to the
// elimination-tree reach of the nonzero rows, numerically nonzero entries compacted on the device and downloaded as CCS
int  chol_device_spsolve(CholDevice* d, int sys, i64 ncols, const i64* Bp, const i64* Bi, const double* Bx, std::vector<i64>& Xp,
                         std::vector<i64>& Xi, std::vector<double>& Xx, CholTimes* times);
void* chol_device_stream(CholDevice* d);      // cudaStream_t
int  chol_device_diag(CholDevice* d, double* diag_host);
int  chol_device_download_L(CholDevice* d, double* L_host);   // raw panel storage, plan.lsize doubles
void chol_device_set_profiling(CholDevice* d, bool on);
int chol_device_download_sign(CholDevice* d, double* sign_host);   // ldl: +-1 per pivot (permuted order)
void chol_device_set_ldl(CholDevice* d, bool on);          // LDL' semantics for sys 2..6 (cholmod.options['supernodal'] = 0)
// level-stepped factorization + front ownership: building blocks of the multi-GPU subtree-to-subcube driver
int  chol_device_set_owned(CholDevice* d, const unsigned char* owned);
int  chol_device_factor_begin(CholDevice* d, const double* val, bool on_device);
int  chol_device_factor_level(CholDevice* d, int level);
int  chol_device_factor_level_phase(CholDevice* d, int level, int phase);
int  chol_device_solve_dist_begin(CholDevice* d, const double* b_dev);
int  chol_device_solve_dist_level(CholDevice* d, int backward, int level);
int  chol_device_solve_dist_end(CholDevice* d, double* x_dev);
int  chol_device_solve_buffers(CholDevice* d, double** T, double** X);
int  chol_device_set_syrk_split(CholDevice* d, const unsigned char* own, const int* lo, const int* hi, const long long* base, double* scratch);
int  chol_device_factor_end(CholDevice* d, i64* minor, CholTimes* times);
int  chol_device_sync(CholDevice* d);
void chol_device_buffers(CholDevice* d, double** L, double** W);
void chol_device_mark_numeric(CholDevice* d, bool numeric);
int  chol_device_factor_enqueue(CholDevice* d, const double* val_dev);   // no read-back, capturable; see chol_device_minor_ptr
const int* chol_device_minor_ptr(CholDevice* d);                         // device word: first non-positive pivot or 0x7fffffff
void chol_device_set_solve_sweeps(CholDevice* d, int mode);   // -1: default; bit 0 forward, bit 1 backward persistent sweep
int  chol_child_lists_check(const CholPlan& P);              // assembly-item / gather-chunk child lists against their definition (test hook)
int  persist_schedule_check(int nr, int nc, int nctas);   // CPU replay of the persistent solve schedules (test hook)
i64  chol_device_workspace_bytes(const CholDevice* d);

}  // namespace b200s
