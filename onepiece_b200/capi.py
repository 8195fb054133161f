"""ctypes binding of include/onepiece_b200.h (the C-ABI of libonepiece_b200.so).

This is the Python-side stub a host application would write; tests and bench.py drive the CUDA path through
it.  There is no fallback: if the shared library is missing the import of this module raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# OPB_LIB_PATH: developer hook to A/B-test alternative builds of the same library
LIB_PATH = os.environ.get("OPB_LIB_PATH") or os.path.join(HERE, "libonepiece_b200.so")

OPB_OK = 0
OPB_ERR_INVALID = -1
OPB_ERR_CUDA = -2
OPB_ERR_CAPACITY = -3
OPB_ERR_UNSUPPORTED = -4
OPB_DEPTH_F32 = 5
OPB_DEPTH_U16 = 2
OPB_STORAGE_F32 = 0
OPB_STORAGE_PACKED16 = 1


class OpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"onepiece_b200 error {code}: {msg}")
        self.code = code


class VolumeDesc(C.Structure):
    _fields_ = [
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("width", C.c_int32), ("height", C.c_int32), ("depth_scale", C.c_float),
        ("voxel_resolution", C.c_float), ("truncation", C.c_float),
        ("near_plane", C.c_float), ("far_plane", C.c_float),
        ("max_cubes", C.c_int32), ("storage", C.c_int32), ("device", C.c_int32),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32), ("shard_axis", C.c_int32),
        ("shard_slab_cubes", C.c_int32),
        ("stream", C.c_void_p),
    ]


class FrameStats(C.Structure):
    _fields_ = [
        ("candidate_cubes", C.c_int32), ("frame_cubes", C.c_int32), ("total_cubes", C.c_int32),
        ("overflow", C.c_int32), ("updated_voxels", C.c_int64),
        ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
        ("select_ms", C.c_float), ("integrate_ms", C.c_float),
    ]


class IcpParams(C.Structure):
    _fields_ = [("max_iteration", C.c_int32), ("threshold", C.c_double), ("scaling", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("T_iterated", C.c_float * 16), ("rmse", C.c_double), ("n_inliers", C.c_size_t),
                ("iterations", C.c_int32), ("status", C.c_int32), ("n_local_pairs", C.c_size_t)]


OPB_ODO_MAX_LEVELS = 6
OPB_ODO_MAX_TRACE = 64


class OdometryDesc(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("depth_scale", C.c_float), ("levels", C.c_int32), ("iterations", C.c_int32 * OPB_ODO_MAX_LEVELS), ("device", C.c_int32),
                ("stream", C.c_void_p)]


class TrackingResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("rmse", C.c_double), ("tracking_success", C.c_int32), ("status", C.c_int32),
                ("n_correspondences", C.c_size_t), ("iterations", C.c_int32), ("corr_per_iteration", C.c_int32 * OPB_ODO_MAX_TRACE),
                ("T_per_iteration", (C.c_float * 16) * OPB_ODO_MAX_TRACE)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m onepiece_b200.build` (onepiece_b200 has no CPU path)")

lib = C.CDLL(LIB_PATH)
_p = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); every function include/onepiece_b200.h declares
SIGNATURES = {
    "opb_last_error": (C.c_char_p, []),
    "opb_device_count": (C.c_int, []),
    "opb_host_alloc": (C.c_int, [C.POINTER(_p), _sz]),
    "opb_host_free": (None, [_p]),
    "opb_free": (None, [_p]),
    "opb_pose_inverse": (None, [_p, _p]),
    "opb_frustum_planes": (None, [C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, _p, _p]),
    "opb_selftest_quotient": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "opb_volume_desc_default": (None, [C.POINTER(VolumeDesc)]),
    "opb_volume_create": (C.c_int, [C.POINTER(VolumeDesc), C.POINTER(_p)]),
    "opb_volume_destroy": (None, [_p]),
    "opb_volume_clear": (C.c_int, [_p]),
    "opb_volume_set_params": (C.c_int, [_p, C.POINTER(VolumeDesc)]),
    "opb_volume_set_profiling": (C.c_int, [_p, C.c_int]),
    "opb_volume_profile_read": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "opb_volume_integrate": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "opb_volume_integrate_async": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "opb_volume_integrate_device": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "opb_volume_synchronize": (C.c_int, [_p]),
    "opb_volume_frame_stats": (C.c_int, [_p, C.POINTER(FrameStats)]),
    "opb_volume_prepare_cubes": (C.c_int, [_p, _p, C.c_int, _p, _p, C.POINTER(_sz)]),
    "opb_volume_compute_bounding": (C.c_int, [_p, _p, C.c_int, _p, _p, _p]),
    "opb_volume_last_frame_cubes": (C.c_int, [_p, _p, C.POINTER(_sz)]),
    "opb_volume_num_cubes": (C.c_int, [_p, C.POINTER(_sz)]),
    "opb_volume_download": (C.c_int, [_p, _p, _p, C.POINTER(_sz)]),
    "opb_volume_upload": (C.c_int, [_p, _p, _p, _sz]),
    "opb_volume_extract_mesh": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_volume_count_mesh": (C.c_int, [_p, C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_mesh_clustering_simplify": (C.c_int, [C.c_int, _p, _p, _sz, _p, _sz, C.c_float, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p),
                                               C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_pointcloud_downsample": (C.c_int, [C.c_int, _p, _p, _p, _sz, C.c_float, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_sz)]),
    "opb_mesh_compute_normals": (C.c_int, [C.c_int, _p, _sz, _p, _sz, _p]),
    "opb_volume_extract_mesh_clustered": (C.c_int, [_p, C.c_float, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_volume_transform": (C.c_int, [_p, _p, C.c_int, C.c_float, C.c_int32, C.POINTER(_p)]),
    "opb_volume_merge": (C.c_int, [_p, _p]),
    "opb_volume_extract_mesh_ordered": (C.c_int, [_p, _p, _sz, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_volume_transform_ordered": (C.c_int, [_p, _p, C.c_int, C.c_float, C.c_int32, _p, _sz, C.POINTER(_p), C.POINTER(_p), C.POINTER(_sz)]),
    "opb_volume_get_desc": (C.c_int, [_p, C.POINTER(VolumeDesc)]),
    "opb_volume_halo_export": (C.c_int, [_p, _p, _p, _sz, C.POINTER(_sz)]),
    "opb_volume_halo_import": (C.c_int, [_p, _p, _p, _sz]),
    "opb_volume_halo_clear": (C.c_int, [_p]),
    "opb_volume_num_ghost_cubes": (C.c_int, [_p, C.POINTER(_sz)]),
    "opb_volume_halo_peer_buffer": (C.c_int, [_p, _sz, C.POINTER(_p), _p]),
    "opb_volume_halo_peer_attach": (C.c_int, [_p, _p, _sz, _p]),
    "opb_volume_halo_exchange_peer": (C.c_int, [_p, C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_volume_halo_exchange_begin": (C.c_int, [_p]),
    "opb_volume_halo_exchange_end": (C.c_int, [_p, C.POINTER(_sz), C.POINTER(_sz)]),
    "opb_volume_frame_ring_buffer": (C.c_int, [_p, C.POINTER(_p), _p]),
    "opb_volume_frame_ring_attach": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "opb_volume_integrate_rows_async": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, _p]),
    "opb_volume_frame_ring_status": (C.c_int, [_p]),
    "opb_prefilter_create": (C.c_int, [C.c_int, _p, C.c_int, C.c_int, C.POINTER(_p)]),
    "opb_prefilter_destroy": (None, [_p]),
    "opb_prefilter_run": (C.c_int, [_p, _p, C.c_int, C.c_float, C.c_int, C.c_double, C.c_double, _p, _p]),
    "opb_prefilter_device_result": (_p, [_p]),
    "opb_prefilter_synchronize": (C.c_int, [_p]),
    "opb_volume_integrate_prefiltered": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, C.c_int, C.c_double, C.c_double]),
    "opb_icp_params_default": (None, [C.POINTER(IcpParams)]),
    "opb_icp_create": (C.c_int, [C.c_int, _p, C.POINTER(_p)]),
    "opb_icp_destroy": (None, [_p]),
    "opb_icp_point_to_plane": (C.c_int, [_p, _p, _sz, _p, _p, _sz, _p, C.POINTER(IcpParams), C.POINTER(IcpResult), _p, _sz]),
    "opb_icp_point_to_point": (C.c_int, [_p, _p, _sz, _p, _sz, _p, C.POINTER(IcpParams), C.POINTER(IcpResult), _p, _sz]),
    "opb_icp_comm_buffer": (C.c_int, [_p, C.POINTER(_p), _p]),
    "opb_ipc_open": (C.c_int, [C.c_int, _p, C.POINTER(_p)]),
    "opb_ipc_close": (C.c_int, [C.c_int, _p]),
    "opb_icp_comm_attach": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "opb_icp_comm_detach": (C.c_int, [_p]),
    "opb_icp_last_search_count": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "opb_icp_estimate_normals": (C.c_int, [_p, _p, _sz, C.c_float, C.c_int, _p]),
    "opb_kdtree_create": (C.c_int, [C.c_int, _p]),
    "opb_kdtree_destroy": (None, [_p]),
    "opb_kdtree_build": (C.c_int, [_p, _p, _sz]),
    "opb_kdtree_search": (C.c_int, [_p, _p, _sz, C.c_int, C.c_int, C.c_float, _p, _p, _p]),
    "opb_kdtree_estimate_normals": (C.c_int, [_p, C.c_float, C.c_int, _p]),
    "opb_kdtree_fpfh": (C.c_int, [_p, _p, C.c_int, C.c_float, _p]),
    "opb_kdtree_dump": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "opb_kdtree_feature_matching": (C.c_int, [_p, _p, _sz, _p, _sz, _p, _p]),
    "opb_ransac_rigid_transformation": (C.c_int, [_p, _p, _p, _sz, C.c_int, C.c_double, C.c_uint64, _p, _p, _p, _p, _p, _p]),
    "opb_simple_ba": (C.c_int, [C.c_int, C.c_int, _p, C.c_int, _p, _p, _p, _p, _p, C.c_int]),
    "opb_simple_ba_from_sums": (C.c_int, [C.c_int, _p, C.c_int, _p, _p, _p]),
    "opb_reject_matches": (C.c_int, [_p, _sz, _p, _sz, _p, _p, _p, C.c_int, C.c_float]),
    "opb_icp_last_launch_count": (C.c_int, [_p, C.POINTER(C.c_int)]),
    "opb_icp_last_search_trace": (C.c_int, [_p, _p, C.c_int]),
    "opb_icp_last_nn": (C.c_int, [_p, _p, _sz]),
    "opb_icp_last_prev_pose": (C.c_int, [_p, _p]),
    "opb_icp_last_stamps": (C.c_int, [_p, _p, C.c_int]),
    "opb_cloud_create": (C.c_int, [C.c_int, _p, C.POINTER(_p)]),
    "opb_cloud_destroy": (None, [_p]),
    "opb_cloud_load_from_depth": (C.c_int, [_p, _p, C.c_int, _p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]),
    "opb_cloud_set_points": (C.c_int, [_p, _p, _sz]),
    "opb_cloud_set_normals": (C.c_int, [_p, _p, _sz]),
    "opb_cloud_size": (C.c_int, [_p, C.POINTER(_sz)]),
    "opb_cloud_download": (C.c_int, [_p, _p, _p]),
    "opb_icp_point_to_plane_clouds": (C.c_int, [_p, _p, _p, _p, C.POINTER(IcpParams), C.POINTER(IcpResult), _p, _sz]),
    "opb_icp_point_to_point_clouds": (C.c_int, [_p, _p, _p, _p, C.POINTER(IcpParams), C.POINTER(IcpResult), _p, _sz]),
    "opb_volume_integrate_cloud": (C.c_int, [_p, _p, _p]),
    "opb_icp_set_profiling": (C.c_int, [_p, C.c_int]),
    "opb_icp_reserve": (C.c_int, [_p, _sz, _sz]),
    "opb_icp_set_async_pairs": (C.c_int, [_p, C.c_int]),
    "opb_icp_wait_pairs": (C.c_int, [_p]),
    "opb_icp_last_timing": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "opb_odometry_desc_default": (None, [C.POINTER(OdometryDesc)]),
    "opb_odometry_create": (C.c_int, [C.POINTER(OdometryDesc), C.POINTER(_p)]),
    "opb_odometry_destroy": (None, [_p]),
    "opb_odometry_set_profiling": (C.c_int, [_p, C.c_int]),
    "opb_odometry_set_loop_form": (C.c_int, [_p, C.c_int]),
    "opb_odometry_last_phases": (C.c_int, [_p, _p]),
    "opb_odometry_last_timing": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "opb_frame_create": (C.c_int, [_p, _p, _p, C.c_int, C.POINTER(_p)]),
    "opb_frame_destroy": (None, [_p]),
    "opb_frame_preprocess": (C.c_int, [_p, _p]),
    "opb_frame_is_preprocessed": (C.c_int, [_p]),
    "opb_frame_image": (C.c_int, [_p, _p, C.c_int, C.c_int, _p]),
    "opb_odometry_dense_tracking_frames": (C.c_int, [_p, _p, _p, _p, C.c_int, C.POINTER(TrackingResult), _p, _sz, _p]),
    "opb_odometry_dense_tracking": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, _p, C.c_int, C.POINTER(TrackingResult), _p, _sz, _p]),
    "opb_odometry_single_iteration": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int, _p, _p, _sz, C.POINTER(_sz)]),
}


def _bind():
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # tests/test_abi.py reports missing symbols explicitly
        fn.restype = res
        fn.argtypes = args


_bind()


def check(rc: int):
    if rc != OPB_OK:
        raise OpbError(rc, lib.opb_last_error().decode(errors="replace"))
