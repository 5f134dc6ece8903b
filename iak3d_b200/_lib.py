"""ctypes binding of libiak3d.so (include/iak3d.h).  No fallback: a missing library or a failing call raises."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libiak3d.so")

OK, NOT_SPD, BAD_PARS = 0, 1, 2
ERR_CUDA, ERR_ARG, ERR_TIMEOUT = -1, -2, -3
MAX_SPL = 12          # IAK3D_MAX_SPL
MAX_SSRE = 8          # IAK3D_MAX_SSRE
FIT_KEEP_CKHICX, FIT_CROSS_SQUARE = 1, 2
MODELX = {"nugget": 0, "spherical": 1, "matern": 2, "wendland": 3}


class Iak3dError(RuntimeError):
    pass


class Pars(C.Structure):
    """struct iak3d_pars (include/iak3d.h)."""
    _fields_ = [
        ("modelx", C.c_int32), ("cmeOpt", C.c_int32), ("sdfdType", C.c_int32 * 3), ("quirk_cd1_unscaled", C.c_int32),
        ("ax", C.c_double), ("nux", C.c_double), ("ad", C.c_double), ("nud", C.c_double),
        ("cx0", C.c_double), ("cx1", C.c_double), ("cd1", C.c_double), ("cxd0", C.c_double), ("cxd1", C.c_double),
        ("cme", C.c_double), ("sdfdPars", (C.c_double * 2) * 3), ("quirk_scale", C.c_double),
        ("sdfdSpl", (C.c_double * MAX_SPL) * 3),
        ("nssre", C.c_int32), ("reserved_", C.c_int32), ("cssre", C.c_double * MAX_SSRE),
    ]


class Model(C.Structure):
    """struct iak3d_model (include/iak3d.h): the switches and bounds readParsIAK3D needs."""
    _fields_ = [
        ("modelx", C.c_int32), ("cmeOpt", C.c_int32), ("prodSum", C.c_int32), ("lnTfmdData", C.c_int32),
        ("sdfdType", C.c_int32 * 3), ("reserved_", C.c_int32), ("nud", C.c_double),
        ("axBnds", C.c_double * 2), ("nuxBnds", C.c_double * 2), ("adBnds", C.c_double * 2), ("tau2Bnds", C.c_double * 2),
        ("sxd1Bnds", C.c_double * 2), ("maxd", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32

_PROTOS = {
    "iak3d_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "iak3d_ctx_destroy": (C.c_int, [_vp]),
    "iak3d_last_error": (C.c_char_p, [_vp]),
    "iak3d_device_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_i64)]),
    "iak3d_launch_count": (_i64, [_vp]),
    "iak3d_dataset_create": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _i32, C.POINTER(_vp)]),
    "iak3d_dataset_destroy": (C.c_int, [_vp]),
    "iak3d_dataset_set_W": (C.c_int, [_vp, _dp, _i32]),
    "iak3d_dataset_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "iak3d_dataset_ids": (C.c_int, [_vp, _ip, _ip]),
    "iak3d_dataset_set_sdfd_spline": (C.c_int, [_vp, _i32, _i32, _dp]),
    "iak3d_dataset_set_lnN_column": (C.c_int, [_vp, _i32]),
    "iak3d_dataset_set_ssre": (C.c_int, [_vp, _i32, _ip, _dp]),
    "iak3d_cov_cross_ssre": (C.c_int, [_vp, C.POINTER(Pars), _i64, _dp, _dp, _ip, _dp, _dp]),
    "iak3d_predict_stage_ssre": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _ip, _dp]),
    "iak3d_factor_build_fetch": (C.c_int, [_vp, C.POINTER(Pars), _i32, _dp, _dp]),
    "iak3d_predict_panel_fetch": (C.c_int, [_vp, _i64, _dp, _i64, C.POINTER(_i64)]),
    "iak3d_cov_diag": (C.c_int, [_vp, C.POINTER(Pars), _dp]),
    "iak3d_fit_create_res": (C.c_int, [_vp, C.POINTER(Pars), _dp, _dp, _dp, C.POINTER(_vp)]),
    "iak3d_predict_fetch_terms": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "iak3d_fit_set_option": (C.c_int, [_vp, _i32, _i32]),
    "iak3d_predict_fetch_CkhiCX": (C.c_int, [_vp, _dp]),
    "iak3d_cov_cross_spl": (C.c_int, [_vp, C.POINTER(Pars), _i64, _dp, _dp, C.POINTER(_dp), _dp]),
    "iak3d_predict_stage_spl": (C.c_int, [_vp, _i64, _dp, _dp, _dp, C.POINTER(_dp)]),
    "iak3d_iacov": (C.c_int, [_vp, _i64, _dp, _i64, _dp, C.c_double, C.c_double, _i32, _dp, _dp, _dp]),
    "iak3d_spatial_cov": (C.c_int, [_vp, _i64, _dp, _i32, C.c_double, C.c_double, C.c_double, _dp]),
    "iak3d_cov_build": (C.c_int, [_vp, C.POINTER(Pars), _dp, _dp]),
    "iak3d_cov_cross": (C.c_int, [_vp, C.POINTER(Pars), _i64, _dp, _dp, _dp]),
    "iak3d_lndet_invCb": (C.c_int, [_vp, _i64, _dp, _i32, _dp, _dp, _dp]),
    "iak3d_nll_terms": (C.c_int, [_vp, C.POINTER(Pars), _dp, _dp, _dp]),
    "iak3d_nll_terms_batch": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(Pars), _dp, _dp, _ip]),
    "iak3d_nll_terms_batch_enqueue": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(Pars)]),
    "iak3d_nll_terms_batch_fetch": (C.c_int, [_vp, _dp, _dp, _ip]),
    "iak3d_nll_terms_batch_wsum": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(Pars), _dp, _ip, _i32, _vp, _dp]),
    "iak3d_nll_terms_batch_wsum_multi": (C.c_int, [_i32, C.POINTER(_vp), _ip, C.POINTER(_vp), C.POINTER(Pars), _dp, _ip, _i32, _i32, _dp]),
    "iak3d_dataset_same_W": (C.c_int, [_vp, _dp, _i32, C.POINTER(_i32)]),
    "iak3d_read_pars": (C.c_int, [C.POINTER(Model), _dp, _i32, C.POINTER(Pars)]),
    "iak3d_nll_vector": (C.c_int, [_vp, C.POINTER(Model), _dp, _i32, _i32, _dp]),
    "iak3d_gradnll_fd": (C.c_int, [_vp, C.POINTER(Model), _dp, _i32, C.c_double, _i32, _dp, _dp]),
    "iak3d_reml_tail": (C.c_int, [_dp, C.c_double, C.c_int64, _i32, _i32, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, C.POINTER(_i32)]),
    "iak3d_nll_reml": (C.c_int, [_vp, C.POINTER(Pars), _i32, _dp, _dp, _dp, _dp, _dp, C.POINTER(_i32)]),
    "iak3d_peer_create": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "iak3d_peer_connect": (C.c_int, [_vp, _vp]),
    "iak3d_peer_allreduce": (C.c_int, [_vp, _vp, _i32, _dp, _dp]),
    "iak3d_peer_destroy": (C.c_int, [_vp]),
    "iak3d_dataset_slot_bytes": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "iak3d_nll_grad": (C.c_int, [_vp, _i32, C.POINTER(Pars), _dp, _i32, _dp, _dp, _dp]),
    "iak3d_fit_create": (C.c_int, [_vp, C.POINTER(Pars), _dp, _dp, C.POINTER(_vp)]),
    "iak3d_fit_destroy": (C.c_int, [_vp]),
    "iak3d_fit_get": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "iak3d_predict": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _dp, _dp]),
    "iak3d_predict_stage": (C.c_int, [_vp, _i64, _dp, _dp, _dp]),
    "iak3d_predict_enqueue": (C.c_int, [_vp]),
    "iak3d_predict_fetch": (C.c_int, [_vp, _dp, _dp]),
    "iak3d_gradhess": (C.c_int, [_vp, C.POINTER(Pars), _i32, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "iak3d_gradhess_shard": (C.c_int, [_vp, C.POINTER(Pars), _i32, _dp, _i32, _i32, _dp, _dp, _dp, _dp, _dp, _dp]),
    "iak3d_sync": (C.c_int, [_vp]),
    "iak3d_timer_start": (C.c_int, [_vp]),
    "iak3d_timer_stop": (C.c_int, [_vp, _dp]),
    "iak3d_flush_l2": (C.c_int, [_vp]),
    "iak3d_fp64_peak": (C.c_int, [_vp, _i32, _dp]),
    "iak3d_mat_create": (C.c_int, [_vp, _i64, _i64, C.POINTER(_vp)]),
    "iak3d_mat_destroy": (C.c_int, [_vp]),
    "iak3d_mat_trim": (C.c_int, [_vp]),
    "iak3d_mat_shape": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "iak3d_mat_set": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _dp, _i64]),
    "iak3d_mat_get": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _dp, _i64]),
    "iak3d_mat_copy": (C.c_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i32, C.c_double, C.c_double]),
    "iak3d_mat_gemm_nt": (C.c_int, [_vp, C.c_double, _vp, _vp, _i32]),
    "iak3d_mat_spd_inverse": (C.c_int, [_vp, _ip]),
    "iak3d_mat_cov_build": (C.c_int, [_vp, C.POINTER(Pars), _vp]),
    "iak3d_mat_coldot": (C.c_int, [_vp, _vp, _dp]),
    "iak3d_design_create": (C.c_int, [_vp, _ip, _i64, _dp, _i64, C.POINTER(_vp)]),
    "iak3d_design_destroy": (C.c_int, [_vp]),
    "iak3d_design_info": (C.c_int, [_vp, _ip, _ip]),
    "iak3d_design_eval": (C.c_int, [_vp, _i64, _dp, _dp, _dp, _dp, _dp]),
    "iak3d_design_eval_lnN": (C.c_int, [_vp, _i64, _dp, _dp, _i32, _ip, _dp, _dp, _dp, _dp]),
    "iak3d_predict_stage_grid": (C.c_int, [_vp, _vp, _i64, _dp, _dp, _i64, _dp, _i64]),
    "iak3d_predict_fetch_X": (C.c_int, [_vp, _dp]),
    "iak3d_points_create": (C.c_int, [_vp, _i64, _dp, C.POINTER(_vp)]),
    "iak3d_points_destroy": (C.c_int, [_vp]),
    "iak3d_points_nearest": (C.c_int, [_vp, _i32, _dp, _ip, C.POINTER(_i64)]),
    "iak3d_voronoi_seed": (C.c_int, [_i64, _dp, _i32, _i32, _ip, _dp]),
    "iak3d_dirichlet_adjacency": (C.c_int, [_i32, _dp, _ip, _i64, C.POINTER(_i64)]),
    "iak3d_last_stage_ms": (C.c_int, [_vp, _dp, _dp, _dp]),
    "iak3d_last_task_profile": (C.c_int, [_vp, C.POINTER(_i64), _i64, C.POINTER(_i64)]),
}
EXPORTS = tuple(sorted(_PROTOS))

_lib = None


def load():
    """Load libiak3d.so and set the prototypes.  Raises if the library is missing (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Iak3dError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)           # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(a, order="F"):
    """Column-major (R native) float64 copy/view."""
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def make_pars(parsBTfmd, modelx, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, quirk_cd1_unscaled=True, quirk_scale=1.0):
    """parsBTfmd (the list readParsIAK3D returns, fitIAK3D.R:1405-1408) -> struct iak3d_pars."""
    P = Pars()
    P.modelx = MODELX[modelx]
    P.cmeOpt = int(cmeOpt)
    types = (int(sdfdType_cd1), int(sdfdType_cxd0), int(sdfdType_cxd1))
    for i, t in enumerate(types):
        P.sdfdType[i] = t
    P.quirk_cd1_unscaled = 1 if quirk_cd1_unscaled else 0
    P.quirk_scale = float(quirk_scale)
    g = lambda k: float(np.asarray(parsBTfmd[k]).reshape(-1)[0]) if np.size(parsBTfmd[k]) else float("nan")
    P.ax, P.nux, P.ad, P.nud = g("ax"), g("nux"), g("ad"), g("nud")
    P.cx0, P.cx1, P.cd1, P.cxd0, P.cxd1, P.cme = g("cx0"), g("cx1"), g("cd1"), g("cxd0"), g("cxd1"), g("cme")
    for i, k in enumerate(("sdfdPars_cd1", "sdfdPars_cxd0", "sdfdPars_cxd1")):
        v = np.asarray(parsBTfmd.get(k, []), float).reshape(-1)
        if types[i] == -1:
            if v.size != 2:
                raise ValueError(f"{k} must hold (tau1, tau2) when its sdfdType is -1")
            P.sdfdPars[i][0], P.sdfdPars[i][1] = float(v[0]), float(v[1])
        elif types[i] > 0:
            if types[i] > MAX_SPL or v.size != types[i]:
                raise ValueError(f"{k} must hold sdfdType = {types[i]} (<= {MAX_SPL}) spline coefficients")
            for e in range(types[i]):
                P.sdfdSpl[i][e] = float(v[e])
    cs = parsBTfmd.get("cssre") if hasattr(parsBTfmd, "get") else None
    if cs is not None and np.size(cs) > 0:          # site-specific random effects (fitIAK3D.R:1413)
        cs = np.asarray(cs, float).reshape(-1)
        if cs.size > MAX_SSRE:
            raise ValueError(f"at most {MAX_SSRE} site-specific random-effect terms")
        P.nssre = cs.size
        for e in range(cs.size):
            P.cssre[e] = float(cs[e])
    return P


def make_model(modelx, nud, sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1, cmeOpt, prodSum, parBnds, lnTfmdData=False):
    """The arguments of readParsIAK3D other than the vector -> struct iak3d_model."""
    M = Model()
    M.modelx = MODELX[modelx]
    M.cmeOpt, M.prodSum, M.lnTfmdData = int(cmeOpt), 1 if prodSum else 0, 1 if lnTfmdData else 0
    for i, t in enumerate((sdfdType_cd1, sdfdType_cxd0, sdfdType_cxd1)):
        M.sdfdType[i] = int(t)
    M.nud = float(nud)
    nan2 = (float("nan"), float("nan"))
    for name in ("axBnds", "nuxBnds", "adBnds", "tau2Bnds", "sxd1Bnds"):
        lo, hi = parBnds.get(name, nan2)
        getattr(M, name)[0], getattr(M, name)[1] = float(lo), float(hi)
    M.maxd = float(parBnds.get("maxd", float("nan")))
    return M


def spl_ptrs(mats):
    """[3] array of double* (or NULL) for the *_spl entry points; `mats` are n x ncol F-ordered arrays or None."""
    arr = (_dp * 3)()
    for i, m in enumerate(mats):
        arr[i] = dptr(m) if m is not None else None
    return arr


class Context:
    """One per GPU (iak3d_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        st = self.lib.iak3d_ctx_create(int(device), C.byref(h))
        if st != OK or not h.value:
            raise Iak3dError(f"iak3d_ctx_create(device={device}) failed with status {st}: an sm_100-class GPU is required "
                             "(there is no CPU fallback)")
        self.h = h
        self.device = int(device)
        self._children = weakref.WeakSet()     # datasets / kept fits: destroyed before the context

    def adopt(self, child):
        self._children.add(child)

    def check(self, st, allow=()):
        if st == OK or st in allow:
            return st
        msg = self.lib.iak3d_last_error(self.h)
        raise Iak3dError(f"libiak3d status {st}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            kids = sorted(list(self._children), key=lambda c: getattr(c, "_close_order", 1))
            for c in kids:
                c.close()
            self.lib.iak3d_ctx_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- measurement helpers -------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.iak3d_launch_count(self.h))

    def sync(self):
        self.check(self.lib.iak3d_sync(self.h))

    def timer_start(self):
        self.check(self.lib.iak3d_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self.check(self.lib.iak3d_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self.check(self.lib.iak3d_flush_l2(self.h))

    def fp64_peak(self, kind):
        t = C.c_double()
        self.check(self.lib.iak3d_fp64_peak(self.h, int(kind), C.byref(t)))
        return t.value

    def last_stage_ms(self):
        ms = (C.c_double * 4)()
        fl, by = C.c_double(), C.c_double()
        self.check(self.lib.iak3d_last_stage_ms(self.h, ms, C.byref(fl), C.byref(by)))
        return list(ms), fl.value, by.value

    def device_info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), _i64()
        self.check(self.lib.iak3d_device_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(sm_count=a.value, cc=(b.value, c.value), total_mem=d.value)


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
