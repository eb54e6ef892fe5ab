"""The C-ABI shared library builds, loads (no GPU needed) and exports every symbol include/sd2b200.h declares with
the argument count the ctypes binding uses."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    h = open(os.path.join(ROOT, 'include', 'sd2b200.h')).read()
    h = re.sub(r'/\*.*?\*/', '', h, flags=re.S)
    out = {}
    for m in re.finditer(r'\n(?:const\s+)?[\w\s\*]+?\b(sd2_\w+)\s*\(([^;{]*?)\)\s*;', h):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ('void', '') else len(args.split(','))
    return out


def test_library_builds_and_exports_every_declared_symbol():
    from diffusion_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    decl = _header_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), f'{name} declared in sd2b200.h but not exported'
    lib.sd2_version.restype = ctypes.c_int
    assert lib.sd2_version() == 100


def test_ctypes_signatures_match_header():
    from diffusion_b200 import _lib
    decl = _header_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_struct_layout_matches_c():
    """sizeof of the ctypes mirrors == sizeof in C (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from diffusion_b200 import _lib
    src = '#include <stdio.h>\n#include "sd2b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(sd2_operand), sizeof(sd2_conv_geom), sizeof(sd2_gemm_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.Operand), ctypes.sizeof(_lib.ConvGeom), ctypes.sizeof(_lib.GemmDesc)]


def test_no_gpu_context_creation_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    from diffusion_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.sd2_ctx_create(0, ctypes.byref(h)) != 0  # no device -> error code, never a silent CPU path
