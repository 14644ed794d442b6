"""Parameter schema of the reference networks, by reference state-dict key name.

The drop-in contract includes weight ingestion by the reference's `state_dict` keys (SURVEY.md section 8b /
Appendix B).  These builders enumerate every tensor (name -> shape) of
  * `ResNet('resnet50')`            -- lib/models/backbones/Resnet.py:6-67,96-135
  * `PoseHighResolutionNet` w32     -- lib/models/backbones/HRnet.py:267-339 + configs/hrnet_w32.yaml:54-93,
                                       instantiated with generate_feat=True, generate_hm=False
  * `RootNetwithRegInt`             -- lib/models/full_net.py:38-192
  * `RootNet`                       -- lib/models/depth_net.py:11-70
They are used to validate incoming checkpoints, to generate synthetic weights, and by the tests to prove the
schema against the real reference (`load_state_dict(strict=True)`).
"""
from __future__ import annotations

from collections import OrderedDict

ROBOTS = {
    # robot_type: (DoF, nkpt, reference_keypoint_id)  -- full_net.py:42-51, configs/*/full.yaml
    "panda": (8, 7, 3),
    "kuka": (7, 8, 3),
    "baxter": (15, 17, 0),
}
DEPTH_DIM = 64
HRNET_CHANNELS = (32, 64, 128, 256)
HRNET_MODULES = {2: 1, 3: 4, 4: 3}   # stage -> NUM_MODULES (hrnet_w32.yaml)
HEAD_CHANNELS = (32, 64, 128, 256)   # HRnet.py:343


def _conv(spec, name, cout, cin, k, bias=False):
    spec[name + ".weight"] = (cout, cin, k, k)
    if bias:
        spec[name + ".bias"] = (cout,)


def _bn(spec, name, c):
    spec[name + ".weight"] = (c,)
    spec[name + ".bias"] = (c,)
    spec[name + ".running_mean"] = (c,)
    spec[name + ".running_var"] = (c,)
    spec[name + ".num_batches_tracked"] = ()


def _bottleneck(spec, name, inplanes, planes, downsample):
    _conv(spec, f"{name}.conv1", planes, inplanes, 1)
    _bn(spec, f"{name}.bn1", planes)
    _conv(spec, f"{name}.conv2", planes, planes, 3)
    _bn(spec, f"{name}.bn2", planes)
    _conv(spec, f"{name}.conv3", planes * 4, planes, 1)
    _bn(spec, f"{name}.bn3", planes * 4)
    if downsample:
        _conv(spec, f"{name}.downsample.0", planes * 4, inplanes, 1)
        _bn(spec, f"{name}.downsample.1", planes * 4)


def _basic(spec, name, c):
    _conv(spec, f"{name}.conv1", c, c, 3)
    _bn(spec, f"{name}.bn1", c)
    _conv(spec, f"{name}.conv2", c, c, 3)
    _bn(spec, f"{name}.bn2", c)


def resnet50_spec(prefix: str = "") -> "OrderedDict[str, tuple]":
    spec = OrderedDict()
    _conv(spec, prefix + "conv1", 64, 3, 7)
    _bn(spec, prefix + "bn1", 64)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3)), start=1):
        for b in range(blocks):
            _bottleneck(spec, f"{prefix}layer{li}.{b}", inplanes, planes, downsample=(b == 0))
            inplanes = planes * 4
    return spec


def hrnet32_spec(prefix: str = "") -> "OrderedDict[str, tuple]":
    spec = OrderedDict()
    C = HRNET_CHANNELS
    _conv(spec, prefix + "conv1", 64, 3, 3)
    _bn(spec, prefix + "bn1", 64)
    _conv(spec, prefix + "conv2", 64, 64, 3)
    _bn(spec, prefix + "bn2", 64)
    for b in range(4):
        _bottleneck(spec, f"{prefix}layer1.{b}", 64 if b == 0 else 256, 64, downsample=(b == 0))
    # transition1: [conv 256->32 s1] , [[conv 256->64 s2]]
    _conv(spec, prefix + "transition1.0.0", C[0], 256, 3)
    _bn(spec, prefix + "transition1.0.1", C[0])
    _conv(spec, prefix + "transition1.1.0.0", C[1], 256, 3)
    _bn(spec, prefix + "transition1.1.0.1", C[1])
    for stage in (2, 3, 4):
        nb = stage
        if stage > 2:  # transition{2,3}: only the new lowest-resolution branch, from the previous last branch
            t = f"{prefix}transition{stage - 1}.{nb - 1}.0"
            _conv(spec, t + ".0", C[nb - 1], C[nb - 2], 3)
            _bn(spec, t + ".1", C[nb - 1])
        for m in range(HRNET_MODULES[stage]):
            mod = f"{prefix}stage{stage}.{m}"
            for br in range(nb):
                for blk in range(4):
                    _basic(spec, f"{mod}.branches.{br}.{blk}", C[br])
            for i in range(nb):          # output branch (multi_scale_output is True everywhere here)
                for j in range(nb):      # input branch
                    f = f"{mod}.fuse_layers.{i}.{j}"
                    if j > i:
                        _conv(spec, f + ".0", C[i], C[j], 1)
                        _bn(spec, f + ".1", C[i])
                    elif j < i:
                        for k in range(i - j):
                            cout = C[i] if k == i - j - 1 else C[j]
                            _conv(spec, f"{f}.{k}.0", cout, C[j], 3)
                            _bn(spec, f"{f}.{k}.1", cout)
    # classification head (HRnet.py:341-388)
    for i in range(4):
        _bottleneck(spec, f"{prefix}incre_modules.{i}.0", C[i], HEAD_CHANNELS[i], downsample=True)
    for i in range(3):
        _conv(spec, f"{prefix}downsamp_modules.{i}.0", HEAD_CHANNELS[i + 1] * 4, HEAD_CHANNELS[i] * 4, 3, bias=True)
        _bn(spec, f"{prefix}downsamp_modules.{i}.1", HEAD_CHANNELS[i + 1] * 4)
    _conv(spec, prefix + "final_feat_layer.0", 2048, 1024, 1, bias=True)
    _bn(spec, prefix + "final_feat_layer.1", 2048)
    return spec


def _linear(spec, name, cout, cin):
    spec[name + ".weight"] = (cout, cin)
    spec[name + ".bias"] = (cout,)


def full_model_spec(robot_type: str) -> "OrderedDict[str, tuple]":
    """`RootNetwithRegInt` with backbone_name=resnet50, rootnet_backbone_name=hrnet32 (configs/*/full.yaml:17-18),
    in module registration order (full_net.py:56-192)."""
    dof, nkpt, _ = ROBOTS[robot_type]
    spec = OrderedDict()
    spec["init_pose"] = (1, dof)
    spec["init_rot"] = (1, 6)
    spec.update(resnet50_spec("reg_backbone."))
    for i, (cin, cout) in enumerate(((2048, 256), (256, 256), (256, 256))):
        spec[f"deconv_layers.{3 * i}.weight"] = (cin, cout, 4, 4)
        _bn(spec, f"deconv_layers.{3 * i + 1}", cout)
    _conv(spec, "final_layer", nkpt * DEPTH_DIM, 256, 1, bias=True)
    _linear(spec, "fc_pose_1", 1024, 2048 + dof)
    _linear(spec, "fc_pose_2", 1024, 1024)
    _linear(spec, "decpose", dof, 1024)
    _linear(spec, "fc_rot_1", 1024, 2048 + 6)
    _linear(spec, "fc_rot_2", 1024, 1024)
    _linear(spec, "decrot", 6, 1024)
    spec.update(hrnet32_spec("rootnet_backbone."))
    _conv(spec, "depth_layer", 1, 2048, 1, bias=True)
    return spec


def depthnet_spec() -> "OrderedDict[str, tuple]":
    """`RootNet('hrnet32')` (depth_net.py:11-70): backbone.* + depth_layer.*"""
    spec = OrderedDict()
    spec.update(hrnet32_spec("backbone."))
    _conv(spec, "depth_layer", 1, 2048, 1, bias=True)
    return spec


def residual_last_bn_names(spec) -> list:
    """BN modules that close a residual branch (BasicBlock.bn2 / Bottleneck.bn3) -- the `gamma_res` knob of the
    synthetic-weight recipe."""
    names = []
    for k in spec:
        if not k.endswith(".weight"):
            continue
        base = k[: -len(".weight")]
        if base.endswith(".bn3"):
            names.append(base)
        elif base.endswith(".bn2") and (base[: -len(".bn2")] + ".conv3.weight") not in spec and ".branches." in base:
            names.append(base)
    return names
