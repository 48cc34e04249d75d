"""The C-ABI library builds, loads on a machine without a GPU, and exports every symbol that
include/pcfd.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from porous_cfd_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pcfd.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pcfd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pcfd.h but not exported by libpcfd_sm100.so'


def test_binding_covers_header(lib):
    from porous_cfd_b200 import _lib
    assert set(declared_symbols()) == set(_lib.SIGNATURES), 'ctypes binding and header disagree'
    assert lib.pcfd_abi_version() == 1


def test_struct_layouts_match_header():
    """sizeof the ctypes mirrors == sizeof the C structs (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from porous_cfd_b200 import _lib
    src = '#include <stdio.h>\n#include "pcfd.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pcfd_intrans_t), sizeof(pcfd_residual_params_t), sizeof(pcfd_dp_peers_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 'sz.c')
        open(c, 'w').write(src)
        exe = os.path.join(td, 'sz')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        a, b, c_ = map(int, subprocess.check_output([exe]).split())
    assert a == ctypes.sizeof(_lib.InTrans) and b == ctypes.sizeof(_lib.ResidualParams) and c_ == ctypes.sizeof(_lib.DpPeers)


def test_no_cpu_fallback():
    """Off-GPU the product path refuses to run instead of silently computing on the CPU."""
    import torch
    from porous_cfd_b200 import _lib, factory, synthetic
    from porous_cfd_b200.dataset.foam_data import FoamData
    spec = synthetic.model_spec('tiny_pigano')
    model = factory.build_model(spec)
    data, labels, domain = synthetic.make_batch(spec['layout'], 1, 8, 8, 2)
    with pytest.raises(_lib.PcfdError):
        model.training_step(FoamData(data, labels, domain), 0)


def test_product_does_not_import_oracle():
    """Nothing under porous-cfd_b200/ may import or reference the oracle."""
    pkg = os.path.join(ROOT, 'porous-cfd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
