"""ctypes mirrors of the structs in include/b200sparse.h"""
import ctypes as C

i64 = C.c_int64

class CholOpts(C.Structure):
    _fields_ = [("supernodal", C.c_int), ("nmethods", C.c_int), ("postorder", C.c_int), ("dbound", C.c_double),
                ("ordering", C.c_int), ("nrelax", C.c_int * 3), ("zrelax", C.c_double * 3), ("block", C.c_int), ("max_merge_cols", C.c_int)]


class CholInfo(C.Structure):
    _fields_ = [(k, i64) for k in ("n", "nsuper", "nnz_L", "nnz_A", "nlevels", "max_front_rows", "max_front_cols",
                                   "factor_bytes", "workspace_bytes")] + \
               [(k, C.c_double) for k in ("flops", "flops_potrf", "flops_trsm", "flops_syrk")] + \
               [("is_numeric", C.c_int), ("minor", i64)] + \
               [(k, C.c_double) for k in ("ms_h2d", "ms_assemble", "ms_factor", "ms_total", "ms_solve", "ms_analyze",
                                          "ms_dense_update", "ms_potrf", "ms_trsm", "ms_extend", "flops_update")] + \
               [("zn", i64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KluInfo(C.Structure):
    _fields_ = [(k, i64) for k in ("n", "nblocks", "nnz_A", "nnz_L", "nnz_U", "nnz_F", "nlevels", "max_block")] + \
               [("flops", C.c_double), ("bytes_per_refactor", i64)] + \
               [(k, C.c_double) for k in ("ms_h2d", "ms_refactor", "ms_solve", "ms_kernel", "ms_dense")] + [("launches", i64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}




class KktInfo(C.Structure):
    _fields_ = [(k, i64) for k in ("n", "ml", "p", "nnz_S", "nterms", "nnz_L", "singular_mode")] + \
               [(k, C.c_double) for k in ("flops", "ms_assemble", "ms_factor", "ms_solve")] + \
               [("Sp", C.POINTER(C.c_int64)), ("Si", C.POINTER(C.c_int64))]


class KktdInfo(C.Structure):
    _fields_ = [(k, i64) for k in ("n", "ml", "p", "launches")] + [(k, C.c_double) for k in ("flops", "ms_factor", "ms_solve")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KluPlanView(C.Structure):
    _fields_ = [(k, i64) for k in ("n", "nlevels", "nslots", "lu_slots", "nnz_A", "nupd", "ndest")] + \
               [(k, C.POINTER(C.c_int64)) for k in ("cbeg", "rowptr", "upd_ptr", "upd_dest")] + \
               [(k, C.POINTER(C.c_int32)) for k in ("udiag_slot", "slot_src", "slot_row", "rowent", "level_ptr", "level_cols",
                                                    "upd_uslot", "upd_lslot", "upd_cnt", "dest", "lslot0", "fslot0")] + \
               [(k, i64) for k in ("nwaves", "nwaves_with_deps", "nbatches", "nsegments", "staged_rows", "npieces", "npiece_users",
                                   "wave_ok", "nearly", "nearly_levels")]
