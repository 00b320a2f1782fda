"""pose_estimation_b200 — B200 (sm_100a) implementation of the pose-refinement hot path of
yumi-crew/pose_estimation: PCL-1.10-semantics VoxelGrid -> NormalEstimation -> ICP.

    pose_estimation_b200.pcl      PCL-shaped host classes over the C ABI of libpe_b200.so
    pose_estimation_b200.testing  synthetic clouds of the BASELINE.json configurations
    csrc/                         the CUDA kernels and the C ABI (include/pe_b200.h)

Importing `pose_estimation_b200.pcl` loads libpe_b200.so and fails loudly if it has not been
built; nothing in this package computes on the CPU.
"""
__version__ = "0.1.0"
