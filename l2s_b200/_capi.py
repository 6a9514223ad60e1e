"""ctypes binding of include/l2s_b200.h.  The CUDA library is the only implementation: if it is missing,
importing the product fails loudly (there is no CPU or PyTorch fallback)."""
import ctypes as C
import os

from . import build as _build

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
c_u32p = C.POINTER(C.c_uint32)
REC_HEADER_WORDS = 4   # L2S_REC_HEADER_WORDS: host record = [npairs | done<<16 | status<<24, makespan, reward, incumbent, pairs...]


class L2SError(RuntimeError):
    """A negative l2s_status returned by the library (or recorded on the device)."""

    def __init__(self, status, where="", instance=None):
        self.status = status
        self.instance = instance
        msg = lib().l2s_status_string(status).decode()
        super().__init__("l2s_b200 %s: %s (status %d%s)" % (
            where, msg, status, "" if instance is None else ", instance %d" % instance))


class StepIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("high", C.c_float), ("fea_norm_const", C.c_float),
                ("reward_type", C.c_int32), ("is_reset", C.c_int32), ("x", C.c_void_p),
                ("edge_index_mc", C.c_void_p), ("makespan", C.c_void_p), ("incumbent", C.c_void_p),
                ("reward", C.c_void_p), ("done", C.c_void_p), ("pairs", C.c_void_p), ("npairs", C.c_void_p),
                ("status", C.c_void_p)]


STATUS = dict(OK=0, INVALID_ARG=-1, CUDA=-2, UNSUPPORTED_SHAPE=-3, ILLEGAL_ACTION=-4, BAD_GRAPH=-5,
              DURATION_RANGE=-6, MACHINE_ROWS=-7, PAIR_OVERFLOW=-8, SHORT_PATH=-9, BAD_SOLUTION=-10)
REWARD = {"yaoxin": 0, "consecutive": 1}
RULE = {"spt": 0, "fdd-divide-mwkr": 1}
POLICY = {"replay": 0, "random": 1, "greedy": 2}

# every symbol include/l2s_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "l2s_status_string": (C.c_char_p, [C.c_int]),
    "l2s_version": (C.c_int, []),
    "l2s_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "l2s_destroy": (C.c_int, [C.c_void_p]),
    "l2s_pmax": (C.c_int, [C.c_void_p]),
    "l2s_edges_mc_per_instance": (C.c_int64, [C.c_void_p]),
    "l2s_itr": (C.c_int64, [C.c_void_p]),
    "l2s_set_itr": (C.c_int, [C.c_void_p, C.c_int64]),
    "l2s_launch_info": (C.c_int, [C.c_void_p, c_i32p, c_i32p, c_i32p]),
    "l2s_load_instances": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "l2s_set_solution": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "l2s_build_initial": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "l2s_get_solution": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "l2s_step": (C.c_int, [C.c_void_p, C.POINTER(StepIO), C.c_void_p]),
    "l2s_step_host": (C.c_int, [C.c_void_p, C.POINTER(StepIO), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "l2s_step_host_begin": (C.c_int, [C.c_void_p, C.POINTER(StepIO), C.c_void_p, C.c_void_p]),
    "l2s_step_host_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "l2s_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "l2s_host_records": (C.c_int, [C.c_void_p, C.POINTER(c_u32p), c_i32p, C.POINTER(c_i32p)]),
    "l2s_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, c_i32p, C.c_int,
                          C.c_void_p, C.c_void_p]),
    "l2s_export_state": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8 + [C.c_void_p]),
    "l2s_poll_error": (C.c_int, [C.c_void_p, c_i32p, C.c_void_p]),
    "l2s_eval_coo_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int64]),
    "l2s_eval_coo": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p]),
}


def lib_path():
    return os.environ.get("L2S_B200_LIB", _build.LIB_PATH)


def lib():
    """Load libl2s_b200.so (once).  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.isfile(path):
            raise RuntimeError(
                "l2s_b200: CUDA library %s not found -- build it with `python -m l2s_b200.build` "
                "(there is no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status, where=""):
    if status < 0:
        raise L2SError(status, where)
    return status


_pylists = None


def pylists():
    """Optional CPython glue (l2s_pylists.so): builds the nested move lists / parses action lists at C speed.
    Returns None if it is not built; callers then use the (slower) numpy route.  Formatting only -- no compute."""
    global _pylists
    if _pylists is None:
        path = _build.PYLISTS_PATH
        if not os.path.isfile(path):
            try:
                _build.build_pylists()
            except Exception:
                _pylists = False
                return None
        try:
            L = C.PyDLL(path)
            L.l2s_build_move_lists.restype = C.py_object
            L.l2s_build_move_lists.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.l2s_parse_actions.restype = C.c_int
            L.l2s_parse_actions.argtypes = [C.py_object, C.c_void_p, C.c_int]
            _pylists = L
        except Exception:
            _pylists = False
    return _pylists or None
