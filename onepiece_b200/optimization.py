"""Python mirror of one_piece::optimization (reference src/Optimization/{Correspondence.h,SimpleBA.h,Optimizer.h}) over the C-ABI:
the pose-graph refinement DenseSlam runs over its submaps, with the per-pair normal-equation sums reduced on the GPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class Correspondence:
    """optimization::Correspondence (Correspondence.h:10-20): frames source_id -> target_id linked by 3-D point pairs, each point
    in its own frame's coordinates"""
    source_id: int
    target_id: int
    source_points: np.ndarray
    target_points: np.ndarray


def SimpleBA(correspondences, camera_poses, max_iteration: int = 5, device: int = 0):
    """optimization::SimpleBA(correspondences, camera_poses, max_iteration) (reference src/Optimization/SimpleBA.cpp:80-157)
    -> the refined poses ([n, 4, 4] float32; pose 0 is fixed).  Like the reference, fewer than three poses are returned as they
    are; an unconnected graph raises."""
    poses = np.ascontiguousarray(camera_poses, np.float32).reshape(-1, 4, 4)
    P = np.ascontiguousarray(poses.transpose(0, 2, 1)).reshape(-1)              # column-major per pose
    sid = np.array([c.source_id for c in correspondences], np.int32)
    tid = np.array([c.target_id for c in correspondences], np.int32)
    a = [np.ascontiguousarray(c.source_points, np.float32).reshape(-1, 3) for c in correspondences]
    b = [np.ascontiguousarray(c.target_points, np.float32).reshape(-1, 3) for c in correspondences]
    if any(len(x) != len(y) for x, y in zip(a, b)):
        raise ValueError("source_points and target_points of a correspondence must pair up")
    off = np.concatenate([[0], np.cumsum([len(x) for x in a])]).astype(np.int64)
    A = np.concatenate(a) if a else np.zeros((0, 3), np.float32)
    B = np.concatenate(b) if b else np.zeros((0, 3), np.float32)

    def ptr(x):
        return x.ctypes.data_as(C.c_void_p)
    capi.check(capi.lib.opb_simple_ba(device, len(poses), ptr(P), len(sid), ptr(sid), ptr(tid), ptr(off), ptr(A), ptr(B), max_iteration))
    return P.reshape(-1, 4, 4).transpose(0, 2, 1).copy()


class Optimizer:
    """optimization::Optimizer (Optimizer.h:10-27); only FastBA is on the path"""

    def FastBA(self, correspondences, poses, max_iteration: int = 5, device: int = 0):
        return SimpleBA(correspondences, poses, max_iteration, device)
