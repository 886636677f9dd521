"""Drop-in check: the REFERENCE's own unit tests (its tests/ directory, unmodified), run against nflows_b200 by aliasing
the import path `nflows` -> `nflows_b200`.  Only runs where /root/reference exists (the build container); the files
selected are the ones that exercise classes on the hot path plus the thin host-side classes kept for
API compatibility.  Out-of-scope classes referenced by those files (NaiveLinear, other piecewise couplings) are deselected."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

REF = "/root/reference"

FILES = ["tests/transforms/base_test.py", "tests/transforms/coupling_test.py", "tests/transforms/splines/rational_quadratic_test.py",
         "tests/transforms/normalization_test.py", "tests/transforms/lu_test.py", "tests/transforms/linear_test.py",
         "tests/transforms/permutations_test.py", "tests/transforms/standard_test.py", "tests/transforms/conv_test.py",
         "tests/transforms/reshape_test.py", "tests/transforms/made_test.py", "tests/flows/base_test.py", "tests/flows/realnvp_test.py",
         "tests/distributions/normal_test.py", "tests/utils/torchutils_test.py"]
DESELECT = ["tests/transforms/linear_test.py::NaiveLinearTest",     # O(D^3) slogdet linear: out of scope
            "tests/transforms/coupling_test.py::UMNNTransformTest"]  # third-party UMNN integrand: out of scope (row 14)
# coupling_test.py lists four piecewise coupling classes in one table; the three with other spline families (SURVEY
# section 2 row 4, out of scope) are dropped from that table, the rational-quadratic one -- the hot path -- stays.
DROP_LINES = ["coupling.PiecewiseLinearCouplingTransform,", "coupling.PiecewiseQuadraticCouplingTransform,",
              "coupling.PiecewiseCubicCouplingTransform,"]

ALIASES = ["transforms", "transforms.base", "transforms.coupling", "transforms.splines", "transforms.splines.rational_quadratic",
           "transforms.normalization", "transforms.linear", "transforms.lu", "transforms.permutations", "transforms.standard",
           "transforms.conv", "transforms.reshape", "transforms.made", "transforms.autoregressive", "transforms.nonlinearities", "distributions", "distributions.base",
           "distributions.normal", "flows", "flows.base", "flows.realnvp", "nn", "nn.nets", "utils", "utils.torchutils",
           "utils.typechecks"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference checkout not present on this box")
def test_reference_unit_tests_pass_against_nflows_b200(tmp_path):
    shutil.copytree(os.path.join(REF, "tests"), str(tmp_path / "tests"))
    ct = tmp_path / "tests" / "transforms" / "coupling_test.py"
    ct.write_text("\n".join(l for l in ct.read_text().splitlines() if l.strip() not in DROP_LINES) + "\n")
    (tmp_path / "conftest.py").write_text(textwrap.dedent("""
        import importlib, sys
        sys.path[:0] = [{root!r}, {shims!r}]
        import nflows_b200
        sys.modules["nflows"] = nflows_b200
        for sub in {aliases!r}:
            sys.modules["nflows." + sub] = importlib.import_module("nflows_b200." + sub)
    """).format(root=ROOT, shims=os.path.join(ROOT, "tests", "_shims"), aliases=ALIASES))
    cmd = [sys.executable, "-m", "pytest", "-q", "--no-header", "-p", "no:cacheprovider"] + FILES
    for d in DESELECT:
        cmd += ["--deselect", d]
    # the reference's tests draw unseeded random inputs against tight eps values: a failing run is repeated once before it counts
    for attempt in range(2):
        out = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
        if out.returncode == 0:
            break
    tail = "\\n".join(out.stdout.splitlines()[-25:])
    assert out.returncode == 0, tail
