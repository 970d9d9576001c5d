"""ctypes binding of libclk.so (include/clk.h).

Every wrapper takes torch CUDA tensors (or raw ints) and passes `data_ptr()`s plus the current
torch CUDA stream through the C ABI.  There is no fallback: a missing library or a non-sm_100
device raises immediately.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLK_LIB_PATH") or os.path.join(_HERE, "libclk.so")  # CLK_LIB_PATH: developer A/B of two builds

CLK_OK = 0
_ERRNAMES = {-1: "CLK_E_BADARG", -2: "CLK_E_UNSUPPORTED_SHAPE", -3: "CLK_E_WORKSPACE", -4: "CLK_E_CUDA",
             -5: "CLK_E_ARCH"}


class ClkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


p, i, ll, f, d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double

# name -> argtypes (the trailing stream argument included), mirrors include/clk.h one to one
SIGNATURES = {
    "clk_nchw_f32_to_nhwc_bf16": [p, p, i, i, i, i, i, p],
    "clk_nhwc_to_nchw_f32": [p, i, p, i, i, i, i, i, p],
    "clk_im2col3x3_stem": [p, p, i, i, i, i, p],
    "clk_stem_conv3x3_fprop": [p, i, p, p, p, p, p, p, p, i, i, i, i, p],
    "clk_pack_w": [p, p, p, i, i, i, i, i, i, i, i, p],
    "clk_unpack_wgrad": [p, p, i, i, i, i, i, f, i, i, p],
    "clk_pack_w_multi": [p, i, i, i, p],
    "clk_unpack_wgrad_multi": [p, i, i, i, p],
    "clk_f64_to_f32_multi": [p, i, p],
    "clk_reduce_partials_multi": [p, i, i, p],
    "clk_conv3x3_fprop": [p, i, p, i, p, p, p, p, p, i, i, i, i, i, p],
    "clk_conv3x3_fprop_eval": [p, i, p, i, p, p, p, p, p, i, i, i, i, i, p],
    "clk_conv3x3_dgrad": [p, i, p, p, i, p, i, i, i, i, p],
    "clk_conv3x3_wgrad": [p, i, p, i, p, i, p, i, i, i, p],
    "clk_conv3x3_wgrad_split": [p, i, p, i, p, i, p, i, i, i, p],
    "clk_gemm_fprop": [p, i, p, p, p, i, i, i, i, p, p, ll, i, p],
    "clk_gemm_fprop_eval": [p, i, p, p, p, i, i, i, p, p, ll, i, p],
    "clk_gemm_wgrad": [p, i, p, i, p, i, i, ll, p],
    "clk_head_argmax_confusion": [p, p, p, p, ll, i, i, i, p, p, p, p],
    "clk_head_loss_bwd": [p, p, p, p, p, p, ll, i, i, i, f, f, f, p, p, p, p, p, p],
    "clk_convT2x2_fprop": [p, p, p, p, i, i, i, i, i, p],
    "clk_convT2x2_dgrad": [p, p, p, i, i, i, i, i, p],
    "clk_convT2x2_wgrad": [p, p, p, i, i, i, i, i, p],
    "clk_bn_stats": [p, p, p, ll, i, p],
    "clk_bn_finalize": [p, p, p, p, p, p, p, p, p, p, i, d, f, f, i, p],
    "clk_bn_apply": [p, p, p, p, ll, i, p],
    "clk_bn_apply_pool": [p, p, p, p, p, p, i, i, i, i, p],
    "clk_maxpool_bwd_add": [p, p, p, p, i, i, i, i, p],
    "clk_maxpool_bwd_add_reduce": [p, p, p, p, p, p, p, i, i, i, i, p],
    "clk_bn_bwd_reduce": [p, p, p, p, ll, i, p],
    "clk_bn_bwd_finalize": [p, p, p, p, p, p, p, p, p, p, i, d, i, i, p],
    "clk_bn_relu_bwd_apply": [p, p, p, p, p, p, p, ll, i, p],
    "clk_channel_sum": [p, p, ll, i, p],
    "clk_f64_to_f32": [p, p, i, i, i, f, i, p],
    "clk_ce_kd_loss": [p, p, p, ll, i, i, f, f, f, p, i, p, p, p],
    "clk_confusion_matrix": [p, p, ll, i, p, p, p],
    "clk_argmax_confusion": [p, p, ll, i, i, p, p, p, p],
    "clk_confusion_matrix_batched": [p, p, i, ll, i, p, p, p],
    "clk_voc_prepare_batch": [p, i, i, i, p, p, p, p],
    "clk_labels_to_rgb": [p, ll, ll, p, p],
    "clk_adam_multi_tensor": [p, p, i, i, f, f, f, f, f, f, f, p, p],
}
PLAIN = {
    "clk_version": ([], i),
    "clk_last_error": ([], C.c_char_p),
    "clk_query_device": ([i], i),
    "clk_set_tuning": ([C.c_char_p, i], i),
    "clk_conv3x3_wgrad_splits": ([i, i, i, i, i], i),
}

_lib = None
_checked_devices = set()


def load():
    """dlopen libclk.so (building is the job of continual_learning_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m continual_learning_b200.build`. "
            "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, argt in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argt
        fn.restype = i
    for name, (argt, rest) in PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = argt
        fn.restype = rest
    _lib = lib
    return lib


def exported_symbols():
    return list(SIGNATURES) + list(PLAIN)


def last_error() -> str:
    return load().clk_last_error().decode()


def check(rc):
    if rc != CLK_OK:
        raise ClkError(rc, last_error())


def ensure_device(dev=None):
    """Fail loudly unless the current CUDA device is sm_100."""
    if not torch.cuda.is_available():
        raise RuntimeError("continual_learning_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    dev = torch.cuda.current_device() if dev is None else dev
    if dev not in _checked_devices:
        check(load().clk_query_device(dev))
        _checked_devices.add(dev)
        # developer override of the library's tuning knobs: CLK_TUNING="pdl=0,conv3_v2=2"
        for item in filter(None, os.environ.get("CLK_TUNING", "").split(",")):
            k, v = item.split("=")
            set_tuning(k.strip(), int(v))
    return dev


def set_tuning(key: str, value: int):
    check(load().clk_set_tuning(key.encode(), int(value)))


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    assert t.is_cuda, "libclk takes device tensors only"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


launch_count = 0      # clk_* kernel launches issued by this process (bench.py reports it)
param_epoch = 0       # bumped by FusedAdam.step: parameters changed behind autograd's version counters
_profile = None       # when a list: (name, start_event, end_event) per launch, for per-kernel timing


def start_profile():
    global _profile
    _profile = []


def stop_profile():
    """returns {entry point: (launches, total_ms)} measured with CUDA events on the launching stream."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec or []:
        n, ms = out.get(name, (0, 0.0))
        out[name] = (n + 1, ms + e0.elapsed_time(e1))
    return out


def call(name, *args):
    """Call a clk_* kernel entry point: tensors -> pointers, appends the current stream, checks status."""
    global launch_count
    lib = load()
    conv = [(_ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*conv, _stream())
        e1.record()
        _profile.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*conv, _stream())
    launch_count += 1
    if rc != CLK_OK:
        raise ClkError(rc, last_error())
