#!/usr/bin/env bash
# Looks on the GPU image for anything that could pin the oracle: OpenCV-contrib's
# surface_matching, PCL / Eigen / FLANN headers, open3d.  Output: gpurun_out/probe_image.txt
mkdir -p gpurun_out
{
  echo "== cv2"; python -c "import cv2; print(cv2.__version__, 'ppf_match_3d', hasattr(cv2,'ppf_match_3d'), 'flann', hasattr(cv2,'flann_Index'))" 2>&1
  echo "== pip"; python -m pip list 2>/dev/null | grep -i -E "open3d|pcl|opencv|cupy|faiss|trimesh|eigen" 
  echo "== pkg-config"; pkg-config --exists pcl_registration-1.10 2>&1; echo "pcl_registration-1.10 rc=$?"; pkg-config --list-all 2>/dev/null | grep -i -E "pcl|eigen|flann|opencv" 
  echo "== find"; find / -xdev \( -iname "icp.h*" -o -name "eigen3" -o -iname "flann*" -o -iname "pcl-1*" -o -name "Eigen" -o -name "kdtree_single_index.h" -o -name "surface_matching*" \) -not -path "/proc/*" 2>/dev/null | head -50
  echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|Thread|Core" 
  echo "== gpu"; nvidia-smi -L
} > gpurun_out/probe_image.txt 2>&1
