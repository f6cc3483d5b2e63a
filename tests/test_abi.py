"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/emperor_b200.h declares; the descriptor layout matches between C and ctypes."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import REPO, golden_cases, load_golden


def _declared_functions():
    text = open(os.path.join(REPO, "include", "emperor_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/emperor_b200.h but not exported"


def test_python_binding_covers_header(built_lib):
    from astroemperor_b200 import _lib
    bound = {s[0] for s in _lib.SYMBOLS}
    assert bound == set(_declared_functions())
    from astroemperor_b200.modelspec import EMP_ABI_VERSION
    assert _lib.lib().emp_abi_version() == EMP_ABI_VERSION == 6


def test_descriptor_layout_matches_c(built_lib, tmp_path):
    """sizeof/offsetof of EmpModelDesc as the C compiler sees them == ctypes mirror."""
    from astroemperor_b200.modelspec import EmpModelDescC
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "emperor_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(EmpModelDesc), '
                   'offsetof(EmpModelDesc, free_to_full), offsetof(EmpModelDesc, full_init), '
                   'offsetof(EmpModelDesc, prior_ops), sizeof(EmpPriorOp));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    D = EmpModelDescC
    assert [int(x) for x in out] == [ctypes.sizeof(D), D.free_to_full.offset, D.full_init.offset,
                                     D.prior_ops.offset, 64]


def test_sweep_struct_layout_matches_c(built_lib, tmp_path):
    """sizeof/offsetof of EmpPtSweep (the whole-sweep argument block) == its ctypes mirror."""
    from astroemperor_b200._lib import EmpPtSweepC as S
    names = ["p", "betas", "perm", "n_acc", "adapt", "adapt_tau", "sweep_counter", "hist_cap", "D", "chain",
             "store_cap", "store_ring", "peer_p", "peer_logp", "logl_all"]
    src = tmp_path / "layout2.c"
    fmt = " ".join(["%zu"] * (len(names) + 1))
    args = ", ".join(f"offsetof(EmpPtSweep, {n})" for n in names)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "emperor_b200.h"\n'
                   f'int main(){{printf("{fmt}\\n", sizeof(EmpPtSweep), {args});return 0;}}\n')
    exe = tmp_path / "layout2"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).decode().split()]
    assert out == [ctypes.sizeof(S)] + [getattr(S, n).offset for n in names]


def test_create_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: on a box without CUDA the engine refuses to construct."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from astroemperor_b200._lib import EmperorB200Error
    from astroemperor_b200.engine import LikelihoodEngine
    g, spec = load_golden("c1_51peg_k1_p0")
    with pytest.raises(EmperorB200Error):
        LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])


@pytest.mark.parametrize("name", golden_cases())
def test_specs_compile(name):
    g, spec = load_golden(name)
    cm = spec.compile()
    assert cm.ndim_free == g["thetas"].shape[1]
    cm.to_c()


def test_unsupported_prior_is_rejected():
    from astroemperor_b200.modelspec import UnsupportedModelError
    g, spec = load_golden("c1_51peg_k1_p0")
    spec.blocks[0].params[0].prior = "Beta"
    with pytest.raises(UnsupportedModelError):
        spec.compile()
