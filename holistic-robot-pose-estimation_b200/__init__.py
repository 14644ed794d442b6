"""horopose_b200 -- B200-native (sm_100a) inference path of HoRoPose behind a C ABI.

Python here is the host-side mirror of the reference's call signatures:
  * `RootNetwithRegInt` / `get_rootNetwithRegInt_model` / `RootNet` / `get_rootnet`  (lib/models/*.py)
  * `URDFRobot`, `point_projection_from_3d[_tensor]`, `HeatmapIntegralPose`, `uvd_to_xyz`, ... (lib/utils/*.py)
All arithmetic runs in libhrp_b200.so (hand-written CUDA); there is no CPU fallback: importing the compute
entry points without the built library raises.
"""
__version__ = "0.1"
