"""Build + ctypes binding of libsd2b200.so (the C ABI declared in include/sd2b200.h).

There is deliberately no fallback: if the shared library cannot be built or loaded, importing the product path
raises.  The structures below mirror include/sd2b200.h field for field.
"""
import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
_ROOT = os.path.dirname(_HERE)
_BUILD = os.path.join(_ROOT, 'build')
LIB_PATH = os.path.join(_HERE, 'libsd2b200.so')
SOURCES = ['api.cu', 'ddp.cu', 'gemm_tc.cu', 'attn.cu', 'k1_noise_sched.cu', 'norm.cu', 'pointwise.cu', 'sampler.cu', 'encoders.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
    '-Wno-deprecated-gpu-targets'
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_ROOT, 'include', 'sd2b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link diffusion_b200/libsd2b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    # every rank of a torchrun job may find the library stale at the same time: one builds, the others wait on the lock and
    # then find it fresh
    import fcntl
    with open(os.path.join(_BUILD, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]

    def one(src):
        obj = os.path.join(_BUILD, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(_CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    cmd = [nvcc, '-shared', '-Wno-deprecated-gpu-targets', '-o', LIB_PATH] + objs + ['-lpthread', '-ldl', '-lrt']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB_PATH


# ---- struct mirrors of include/sd2b200.h ------------------------------------------------------------------
class Operand(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('mn_major', C.c_int), ('cols', C.c_int), ('rows', C.c_int), ('ld', C.c_longlong),
                ('nb0', C.c_int), ('nb1', C.c_int), ('bs0', C.c_longlong), ('bs1', C.c_longlong)]


class ConvGeom(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('n_planes', C.c_int), ('H', C.c_int), ('W', C.c_int), ('C', C.c_int),
                ('ldc', C.c_longlong), ('ntaps', C.c_int), ('dh', C.c_int * 9), ('dw', C.c_int * 9),
                ('dn', C.c_int * 9), ('wtap', C.c_int * 9)]


class GemmDesc(C.Structure):
    _fields_ = [('kind', C.c_int), ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('batch', C.c_int),
                ('A', Operand), ('B', Operand), ('conv', ConvGeom), ('out_mode', C.c_int), ('out', C.c_void_p),
                ('ldo', C.c_longlong), ('out_nb0', C.c_int), ('out_bs0', C.c_longlong), ('out_bs1', C.c_longlong),
                ('residual', C.c_void_p), ('ldr', C.c_longlong), ('bias', C.c_void_p), ('rowbias', C.c_void_p),
                ('rows_per_group', C.c_int), ('ld_rowbias', C.c_longlong), ('alpha', C.c_float),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_longlong), ('max_splits', C.c_int), ('force_bn', C.c_int),
                ('force_splits', C.c_int), ('gn_partial', C.c_void_p), ('gn_slab', C.c_int),
                ('mse_target', C.c_void_p), ('mse_dpred8', C.c_void_p), ('mse_acc', C.c_void_p), ('mse_dtype', C.c_int),
                ('mse_hw', C.c_int)]


GEMM_PLAIN, GEMM_CONV, GEMM_CONV_WGRAD = 0, 1, 2
OUT_BF16, OUT_F32, OUT_F32_ACCUM = 0, 1, 2
DT_F32, DT_BF16, DT_F16 = 0, 1, 2

_vp, _i, _ll, _f, _u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64

# name -> (restype, argtypes); ctx is always the first argument, stream always the last
SIGNATURES = {
    'sd2_version': (_i, []),
    'sd2_ctx_create': (_i, [_i, C.POINTER(_vp)]),
    'sd2_ctx_destroy': (_i, [_vp]),
    'sd2_last_error': (C.c_char_p, [_vp]),
    'sd2_num_sms': (_i, [_vp]),
    'sd2_launch_count': (_ll, [_vp]),
    'sd2_noise_sched_fwd': (_i, [_vp, _u64, _u64, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i,
                                 C.POINTER(_u64), _vp]),
    'sd2_timestep_embedding': (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    'sd2_nchw4_to_nhwc8': (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    'sd2_nhwc8_to_nchw4': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'sd2_scale_by_scalar': (_i, [_vp, _vp, _ll, _vp, _vp]),
    'sd2_fill_f32': (_i, [_vp, _vp, _ll, _f, _vp]),
    'sd2_gemm': (_i, [_vp, C.POINTER(GemmDesc), _vp]),
    'sd2_groupnorm_ws_floats': (_ll, [_i, _i]),
    'sd2_groupnorm_fwd': (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'sd2_concat_stats': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'sd2_groupnorm_fwd_fused': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'sd2_groupnorm_bwd': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp, _ll, _vp, _vp, _vp, _i, _i, _i,
                               _i, _i, _vp, _vp, _vp, _vp]),
    'sd2_layernorm_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    'sd2_layernorm_ws_floats': (_ll, [_ll, _i]),
    'sd2_layernorm_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    'sd2_softmax_fwd': (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _i, _vp]),
    'sd2_softmax_bwd': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _ll, _i, _f, _vp]),
    'sd2_attn_fwd': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    'sd2_attn_bwd_ws_bytes': (_ll, [_i, _i, _i]),
    'sd2_attn_bwd': (_i, [_vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _i, _i,
                          _i, _i, _i, _f, _vp]),
    'sd2_geglu_fwd': (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    'sd2_geglu_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    'sd2_silu_fwd': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'sd2_silu_bwd': (_i, [_vp, _vp, _vp, _vp, _ll, _vp]),
    'sd2_axpby': (_i, [_vp, _vp, _f, _vp, _f, _vp, _ll, _vp]),
    'sd2_copy2d': (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _i, _i, _vp]),
    'sd2_upsample2x_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'sd2_upconv_weff_build': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'sd2_upconv_wgrad_scatter': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'sd2_upsample2x_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'sd2_phase_split': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'sd2_phase_merge': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'sd2_colsum': (_i, [_vp, _vp, _ll, _vp, _ll, _i, _ll, _i, _i, _vp]),
    'sd2_cast_f32_to_bf16': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'sd2_pad_cast_rows': (_i, [_vp, _vp, _i, _vp, _i, _ll, _vp]),
    'sd2_unpad_accum_rows': (_i, [_vp, _vp, _i, _vp, _i, _ll, _i, _vp]),
    'sd2_mse_head': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _f, _i, _i, _i, _vp]),
    'sd2_adamw_step': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _i, _vp]),
    'sd2_cfg_ddim_step': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    'sd2_ema_update': (_i, [_vp, _vp, _vp, _ll, _f, _f, _vp]),
    'sd2_cast_to_bf16': (_i, [_vp, _vp, _i, _vp, _ll, _vp]),
    'sd2_wire_gather': (_i, [_vp, _i, _ll, _vp]),
    'sd2_nchw_to_nhwc8': (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    'sd2_nhwc8_to_nchw': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _f, _vp]),
    'sd2_vae_sample': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'sd2_embed_tokens': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    'sd2_softmax_causal_fwd': (_i, [_vp, _vp, _ll, _vp, _ll, _ll, _i, _i, _vp]),
    'sd2_gelu_fwd': (_i, [_vp, _vp, _vp, _ll, _vp]),
    'sd2_pixel_linear8': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    'sd2_ddp_unique_id': (_i, [_vp, _vp]),
    'sd2_ddp_init': (_i, [_vp, _vp, _i, _i]),
    'sd2_ddp_world': (_i, [_vp]),
    'sd2_ddp_allreduce_bucket': (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    'sd2_ddp_destroy': (_i, [_vp]),
    'sd2_workspace_bytes': (_ll, [_ll, _ll, _i]),
}

_lib = None


def load(auto_build=True):
    """Load (building first if needed) the C-ABI library; raises if unavailable - there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get('SD2_LIB') or LIB_PATH  # SD2_LIB: another build of the same ABI (kernel A/B measurements on one box)
    if path == LIB_PATH and auto_build and _stale():
        build()
    if not os.path.exists(path):
        raise RuntimeError(f'{path} is missing: run `python -c "import __graft_entry__ as g; g.build()"`')
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name):
            raise RuntimeError(f'{LIB_PATH} does not export {name}')
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DryLib:
    """Host-logic test double (SD2_DRY_RUN=1, used by the CPU test-suite only): every entry point type-checks its
    arguments against SIGNATURES and counts the call; nothing is computed.  Never used by the product path."""

    def __init__(self):
        self.calls = {}
        for name, (res, args) in SIGNATURES.items():
            setattr(self, name, self._make(name, res, args))

    def _make(self, name, res, args):

        def fn(*a):
            if len(a) != len(args):
                raise TypeError(f'{name}: expected {len(args)} arguments, got {len(a)}')
            for i, (t, v) in enumerate(zip(args, a)):
                try:
                    if hasattr(t, 'from_param'):
                        t.from_param(v)
                except Exception as e:  # noqa: BLE001
                    raise TypeError(f'{name}: argument {i} ({v!r}) is not a {t}') from e
            self.calls[name] = self.calls.get(name, 0) + 1
            if name == 'sd2_groupnorm_ws_floats':
                return a[0] * 64 * a[1] * 3 + a[0] * 128
            if name == 'sd2_attn_bwd_ws_bytes':
                return 8 * a[0] * a[2] * a[1] * 64 * 4 + a[0] * a[1] * ((a[2] + 127) // 128) * 256 * 4
            if name == 'sd2_layernorm_ws_floats':
                return 148 * 4 * a[1] * 2
            if name == 'sd2_last_error':
                return b'dry run'
            if name == 'sd2_num_sms':
                return 148
            return 0

        return fn
