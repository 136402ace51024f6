"""CPU: the C-ABI library loads and exports exactly what include/fsg_dense.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fsg_dense.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"FSG_API\s+[\w\s\*]+?\b(fsg_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from full_scale_gambler_for_object_detection_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.LIB_PATH


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) == 38, syms
    for s in ("fsg_pairwise_iou", "fsg_matcher", "fsg_match_anchors", "fsg_box2box_get_deltas",
              "fsg_box2box_apply_deltas", "fsg_loss_main", "fsg_loss_post", "fsg_nms", "fsg_detect",
              "fsg_permute_level", "fsg_loss_main_levels", "fsg_grid_anchors", "fsg_postprocess_boxes"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (fsg_\w+)", out))
    assert set(declared_symbols()) == exported


def test_ctypes_prototypes_cover_the_header(lib_path):
    from full_scale_gambler_for_object_detection_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    L = _lib.lib()
    assert L.fsg_abi_version() == _lib.ABI_VERSION
    assert L.fsg_status_string(0) == b"ok"
    assert b"workspace" in L.fsg_status_string(2)
    # pure host-side queries (no device work)
    assert L.fsg_match_workspace_bytes(2, 1000, 16) > 2 * 1000 * 8
    assert L.fsg_loss_main_workspace_bytes(2, 1000, 80) >= 16
    assert L.fsg_detect_workspace_bytes(2, 1000, 80, 5, 1000) > 0
    assert L.fsg_nms_workspace_bytes(100) > 0
    assert ctypes.sizeof(_lib.LossParams) == 64
    assert ctypes.sizeof(_lib.PeerCtx) == 96
    assert ctypes.sizeof(_lib.StepIO) == 19 * 8 and ctypes.sizeof(_lib.MatchConfig) == 64


def test_header_is_plain_c():
    """The boundary compiles as C (no torch / C++ types in the signatures)."""
    src = "#include \"fsg_dense.h\"\nint main(void){ fsg_loss_params p; (void)p; return FSG_ABI_VERSION - 1; }\n"
    exe = "/tmp/fsg_header_check"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-",
                        "-o", exe], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing in the product imports the oracle."""
    import torch
    from full_scale_gambler_for_object_detection_b200 import _lib

    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))
    pkg = os.path.join(ROOT, "full_scale_gambler_for_object_detection_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_argument_validation_needs_no_device(lib_path):
    """Every entry point validates its arguments before touching the device: bad calls come back with a status
    (1 = invalid argument, 2 = workspace, 3 = unsupported shape) instead of launching anything."""
    import ctypes as C
    from full_scale_gambler_for_object_detection_b200 import _lib

    L = _lib.lib()
    INVALID, WORKSPACE, UNSUPPORTED = 1, 2, 3
    null = None
    assert L.fsg_pairwise_iou(null, -1, null, 1, null, null) == INVALID
    assert L.fsg_nms(null, null, null, -1, 0.5, null, null, null, 0, null) == INVALID
    assert L.fsg_nms_workspace_bytes(300000) == 0                       # beyond the general-n cap
    assert L.fsg_nms_workspace_bytes(20000) > 20000 * (20000 // 64) * 8  # bit matrix for the large path
    assert L.fsg_box2box_get_deltas(null, null, -1, null, null, null) == INVALID
    assert L.fsg_permute_level(null, null, 1, 0, 4, 0, 0, 0, null) == INVALID
    assert L.fsg_scale_inplace(null, -1, null, 1.0, null) == INVALID
    assert L.fsg_postprocess_boxes(null, -1, 1.0, 1.0, 1.0, 1.0, null, null, null) == INVALID
    assert L.fsg_score_filter(null, 3, null, 10, 80, 1.0, 1.0, 0.5, null, null, null, null, C.c_void_p(16), null) == INVALID
    lv = (_lib.AnchorLevel * 1)()
    lv[0].H, lv[0].W, lv[0].stride, lv[0].A = 4, 4, 8, 17                # more cell anchors than the table holds
    assert L.fsg_grid_anchors(lv, 1, null, 4 * 4 * 17, null) == INVALID
    lv[0].A = 3
    assert L.fsg_grid_anchors(lv, 1, null, 5, null) == INVALID           # R does not match the grid
    sizes = _lib.host_i64([100, 50])
    assert L.fsg_rpn_proposals_workspace_bytes(2, sizes, 2, 9000, 1000) > 0      # k clipped to the level sizes
    big = _lib.host_i64([20000, 20000])
    assert L.fsg_rpn_proposals_workspace_bytes(2, big, 2, 12000, 2000) > 0       # general-n NMS path (C4 RPN setting)
    assert L.fsg_rpn_proposals_workspace_bytes(2, big, 2, 17000, 2000) == 0      # pre_nms_topk > 16384 per level
    hl = (_lib.HeadLevel * 1)()
    hl[0].H, hl[0].W = 0, 4
    assert L.fsg_loss_main_levels_workspace_bytes(2, hl, 1, 3) == 0
    p = _lib.LossParams()
    p.num_classes = 80
    assert L.fsg_loss_main(null, null, null, null, 0, null, null, null, null, null, null, 1, 10, C.byref(p), null, null,
                           null, null, null, null, null, 0, null) == INVALID
    assert L.fsg_detect(null, null, null, 0, 1, 10, 80, null, 1, 0.05, 1000, 0.5, 100, null, 4.0, null, null, null,
                        null, null, null, null, null, null, null, null, 0, null) == INVALID
    assert b"invalid" in L.fsg_status_string(INVALID).lower() or b"argument" in L.fsg_status_string(INVALID).lower()
    assert L.fsg_status_string(UNSUPPORTED) and L.fsg_status_string(WORKSPACE)


def test_no_dependent_loads_before_pdl_wait(lib_path):
    """Kernels launched under programmatic dependent launch may read only the step's INPUTS before
    griddepcontrol.wait (SASS: ACQBULK).  nvcc hoists loads through `const __restrict__` pointers above the wait
    (common.cuh: produced_by_dependency), which reads the preceding kernel's results while it is still running.
    The built library is disassembled: the only kernel with global loads in front of its wait is pass B of the
    matcher -- gt_offsets[n], gt_offsets[n + 1], its GT box and GT class (all inputs) -- and nothing stores there."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "pdl_audit.py")
    out = subprocess.run(["python", script, lib_path], capture_output=True, text=True, check=True).stdout
    kernels = [l for l in out.splitlines() if re.match(r"\s*\d+  ", l)]
    assert len(kernels) >= 10, out[-2000:]           # every kernel of the PDL chains is in the list
    offenders = {l.split(None, 1)[1].split("(")[0]: int(l.split()[0]) for l in kernels if int(l.split()[0]) > 0}
    assert list(offenders) == ["void fsg::match_pass_b_kernel<4>"], offenders
    pre = [l.strip() for l in out.splitlines() if l.startswith("        ")]
    assert len(pre) == 4 and all(re.match(r"(@!?P\d+ )?LDG\.E(\.\d+)?\.CONSTANT ", l) for l in pre), pre
