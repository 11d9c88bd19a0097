"""The jax.ffi boundary exists as source (vbsky_b200/csrc/ffi/vbsky_ffi.cc, vbsky_b200/jax_ffi.py) but cannot be built in
this image (no jaxlib headers).  These CPU tests keep it honest: the translation unit is type-checked by the host
compiler against the real include/vbsky_b200.h and a declaration-only stand-in for xla/ffi/api/ffi.h whose
XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts, like the real one, that every handler is invocable with exactly the decoded
types of its binding -- so a C entry point called with the wrong arity or pointer types, or a binding that drifts from
its handler, fails here.  The Python module's targets are checked against the handler symbols."""
import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TU = os.path.join(ROOT, "vbsky_b200", "csrc", "ffi", "vbsky_ffi.cc")
CUDA_INC = "/usr/local/cuda/include"


def _handlers():
    src = open(TU).read()
    return dict(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(\w+)\s*,\s*(\w+)\s*,", src)), src


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="needs g++ and the CUDA headers")
def test_ffi_translation_unit_type_checks_against_the_header():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "ffi_stub"), "-I", CUDA_INC, TU]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]


def test_every_replaced_reference_callee_has_a_handler_calling_the_c_abi():
    handlers, src = _handlers()
    want = {"VbskyPrune": "vbsky_prune_value_and_grad", "VbskyBdsky": "vbsky_bdsky_value_and_grad",
            "VbskyHeightsForward": "vbsky_heights_forward", "VbskyHeightsBackward": "vbsky_heights_backward",
            "VbskyLoglik": "vbsky_loglik_value_and_grad", "VbskyReduceElbo": "vbsky_reduce_elbo"}
    assert set(want) <= set(handlers)
    header = open(os.path.join(ROOT, "include", "vbsky_b200.h")).read()
    for sym, cfn in want.items():
        impl = handlers[sym]
        body = src[src.index(f"ffi::Error {impl}("):]
        body = body[: body.index("\n}\n")]
        assert re.search(rf"\b{cfn}\(", body), f"{impl} does not call {cfn}"
        assert re.search(rf"\b{cfn}\s*\(", header), f"{cfn} is not declared in the header"
        assert "stream" in body.split(cfn, 1)[1], f"{impl} must pass XLA's stream to {cfn}"


def test_python_module_registers_exactly_the_handler_symbols():
    handlers, _ = _handlers()
    tree = ast.parse(open(os.path.join(ROOT, "vbsky_b200", "jax_ffi.py")).read())
    targets = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "TARGETS":
            targets = ast.literal_eval(node.value)
    assert targets and set(targets.values()) == set(handlers)
    src = open(os.path.join(ROOT, "vbsky_b200", "jax_ffi.py")).read()
    for name in targets:
        assert f'"{name}"' in src.split("TARGETS", 2)[-1], f"target {name} is registered but never called"
    for fn in ("data_loglik", "tree_loglik", "heights", "loglik"):
        assert f"{fn}.defvjp(" in src  # one custom_vjp per replaced callee
    # the package never imports the JAX module (JAX is optional)
    for f in os.listdir(os.path.join(ROOT, "vbsky_b200")):
        if f.endswith(".py") and f != "jax_ffi.py":
            assert "jax_ffi" not in open(os.path.join(ROOT, "vbsky_b200", f)).read()
