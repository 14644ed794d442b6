"""Host-side logic: URDF table builder, synthetic recipe determinism, C-ABI library surface (no GPU)."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import horopose_b200  # noqa: E402,F401
from horopose_b200 import arch, synth, urdf  # noqa: E402
from oracle import horopose_oracle as O  # noqa: E402


@pytest.mark.parametrize("rt", ["panda", "kuka", "baxter"])
def test_urdf_table_matches_oracle(rt):
    tree = urdf.load_urdf(synth.URDF_PATHS[rt])
    orob = O.OracleRobot(rt, str(synth.URDF_PATHS[rt]))
    assert tree.actuated_joint_names == [j["name"] for j in orob.actuated]
    assert set(tree.link_names) == set(orob.links)
    for i, name in enumerate(tree.link_names):
        p = tree.parent[i]
        assert p < i  # parents before children
        if p < 0:
            assert name == orob.base
            continue
        j = orob.joint_of_child[name]
        assert tree.link_names[p] == j["parent"]
        np.testing.assert_array_equal(tree.origin[i], j["origin"])
        np.testing.assert_array_equal(tree.axis[i], j["axis"])
    from horopose_b200.tables import JOINT_NAMES
    assert tree.actuated_joint_names == JOINT_NAMES[rt]  # stable depth sort restores the authors' order


def test_synth_is_deterministic():
    a = synth.uniform01("abc", 1000, 3)
    b = synth.uniform01("abc", 1000, 3)
    assert np.array_equal(a, b) and a.min() >= 0 and a.max() < 1
    # known-answer: guards against a numpy PCG64 stream change between hosts
    assert a[:3].tolist() == pytest.approx(synth.uniform01("abc", 3, 3).tolist(), abs=0)
    sd = synth.full_state_dict("panda")
    assert float(sd["depth_layer.bias"]) == 2.0
    assert sd["reg_backbone.bn1.running_var"].min() > 0


def test_arch_counts():
    # 2308 keys (Panda), 1956 for the standalone RootNet; parameter totals
    spec = arch.full_model_spec("panda")
    assert len(spec) == 2308
    assert len(arch.depthnet_spec()) == 1956

    def nparams(spec):
        return sum(int(np.prod(s)) for k, s in spec.items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked", "init_pose", "init_rot")))
    assert nparams(spec) == 79_620_431
    assert nparams(arch.full_model_spec("kuka")) == 79_634_830
    assert nparams(arch.full_model_spec("baxter")) == 79_799_254
    assert nparams(arch.depthnet_spec()) == 39_185_729


def test_c_abi_exports_every_declared_symbol():
    from horopose_b200 import _lib
    lib = _lib.lib()
    header = (ROOT / "include" / "hrp.h").read_text()
    names = set(re.findall(r"\b(hrp_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 10
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/hrp.h but not exported by libhrp_b200.so"
    assert b"sm_100a" in lib.hrp_version()
    # argument validation without a GPU: a NULL descriptor is rejected with a message, not a crash
    n = ctypes.c_int64(0)
    assert lib.hrp_conv_packed_weight_elems(None, ctypes.byref(n)) == -1
    assert b"null" in lib.hrp_last_error()


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Rows f1-f4: every new C entry point rejects bad arguments (status -1 + message) before touching CUDA."""
    from horopose_b200 import _lib
    from horopose_b200.metrics import MetricsArgs
    from horopose_b200.preprocess import CropArgs
    lib = _lib.lib()
    assert lib.hrp_crop_resize(None, None) == -1 and b"null" in lib.hrp_last_error()
    a = CropArgs(B=1, frame_h=480, frame_w=640, out_size=250)            # not a multiple of 4
    assert lib.hrp_crop_resize(ctypes.byref(a), None) == -1 and b"out_size" in lib.hrp_last_error()
    assert lib.hrp_metrics_batch(None, None) == -1
    m = MetricsArgs(B=4, nkpt=40, dof=7, ref_kpt=0)                      # more keypoints than a warp has lanes
    assert lib.hrp_metrics_batch(ctypes.byref(m), None) == -1 and b"nkpt" in lib.hrp_last_error()
    n = ctypes.c_int64(0)
    assert lib.hrp_metrics_workspace_bytes(512, 17, 15, ctypes.byref(n)) == 0 and n.value == (3 * 17 + 15) * 512 * 4
    assert lib.hrp_metrics_summary(None, None, ctypes.c_int64(10), None, None) == -1
    one = ctypes.c_float(0.0)
    p = ctypes.byref(one)
    assert lib.hrp_pnp(p, p, p, 0, 4, 5, p, None, None) == -1 and b"6..64" in lib.hrp_last_error()   # < 6 points
    assert lib.hrp_head_backward_heatmap(None, p, p, p, ctypes.c_int64(0), 1, 7, 3, 1, 1, p, None) == -1
    assert lib.hrp_head_backward_heatmap(p, p, p, p, ctypes.c_int64(8), 1, 7, 3, 1, 1, p, None) == -1
    assert b"workspace" in lib.hrp_last_error()


def test_activation_arena_planning_runs_on_the_host():
    """The two-pass plan builder's first pass (shapes + liveness, no device work) through the C ABI: the liveness-aliased
    arena of a single-lane 512-image plan is a fraction of one-buffer-per-tensor, the planner's own overlap check passes,
    and without aliasing the arena is exactly the sum of the tensors."""
    import numpy as np
    from horopose_b200 import _lib, synth
    from horopose_b200.models import MODEL_DEPTHNET, MODEL_FULL, ModelDesc
    lib = _lib.lib()
    for kind, sd, nk, dof in ((MODEL_FULL, synth.full_state_dict("kuka", with_bn_stats=False), 8, 7),
                              (MODEL_DEPTHNET, synth.depthnet_state_dict(with_bn_stats=False), 0, 0)):
        desc = ModelDesc(kind, dof, nk, 3 if nk else 0, 4, 1, 256.0, 1.3, 512, 1)
        h = ctypes.c_void_p(0)
        assert lib.hrp_model_create(ctypes.byref(desc), ctypes.byref(h)) == 0, lib.hrp_last_error()
        try:
            for k, v in sd.items():
                if v.dtype.is_floating_point and v.dim() == 4:   # the planner only asks which downsample convs exist
                    a = np.ascontiguousarray(v.float().numpy())
                    shape = (ctypes.c_int64 * a.ndim)(*a.shape)
                    assert lib.hrp_model_set_tensor(h, k.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), shape,
                                                    a.ndim) == 0
            res = {}
            for alias in (1, 0):
                arena, total, nt, no = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_int32(0)
                rc = lib.hrp_model_plan_memory(h, 512, alias, ctypes.byref(arena), ctypes.byref(total), ctypes.byref(nt),
                                               ctypes.byref(no))
                assert rc == 0, lib.hrp_last_error()
                res[alias] = (arena.value, total.value, nt.value, no.value)
            assert res[0][0] == res[0][1] == res[1][1]            # no aliasing: arena == sum of the tensors
            assert res[1][0] < 0.35 * res[1][1], res              # aliased: a fraction of it
            if kind == MODEL_FULL:
                assert 40e9 < res[1][1] < 60e9 and res[1][0] < 10e9, res   # ~47 GB of tensors at 512 images -> ~7 GB arena
                assert 370 <= res[1][3] <= 400, res                        # conv / pool ops of the program
        finally:
            lib.hrp_model_destroy(h)
