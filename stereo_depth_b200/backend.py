"""The reference's backend plugin interface for the "cuda" stereo-matching backend.

  StereoMatching             <- src/python/pipeline/depth/stereo_matching.py:6-10
  CudaStereoMatchingBackend  <- src/python/pipeline/depth/cuda_stereo_matching_backend.py:7-17
"""
from abc import ABC, abstractmethod

import torch

from . import cuda_depth


class StereoMatching(ABC):

    @abstractmethod
    def process(self, left_image: torch.Tensor, right_image: torch.Tensor) -> torch.Tensor:
        pass


class CudaStereoMatchingBackend(StereoMatching):

    def __init__(self, configuration: "cuda_depth.StereoMatchingConfiguration" = None, frames_per_launch: int = 0):
        if configuration is None:
            configuration = cuda_depth.StereoMatchingConfiguration()
        self._stereo_algo = cuda_depth.StereoMatching(configuration, frames_per_launch=frames_per_launch)

    @staticmethod
    def _prepare(image: torch.Tensor) -> torch.Tensor:
        # The reference does `.cuda().float().contiguous()`.  uint8 stays uint8 here: the kernels
        # convert on load (exact), which saves the cast kernel and 3/4 of the H2D bytes.
        image = image.cuda()
        if image.dtype not in (torch.uint8, torch.float32):
            image = image.float()
        return image.contiguous()

    def process(self, left_image: torch.Tensor, right_image: torch.Tensor) -> torch.Tensor:
        left_gpu, right_gpu = self._prepare(left_image), self._prepare(right_image)
        if left_gpu.dtype != right_gpu.dtype:
            left_gpu, right_gpu = left_gpu.float(), right_gpu.float()
        return self._stereo_algo.compute_disparity_map(left_gpu, right_gpu)

    def process_batch(self, left_images: torch.Tensor, right_images: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """[N,3,H,W] pairs -> [N,H,W].  CPU inputs take the pipelined host path and return a pinned CPU tensor;
        CUDA inputs stay on the device."""
        if not left_images.is_cuda and not right_images.is_cuda:
            if left_images.dtype not in (torch.uint8, torch.float32):
                left_images, right_images = left_images.float(), right_images.float()
            return self._stereo_algo.compute_disparity_host(left_images.contiguous(), right_images.contiguous(), out)
        left_gpu, right_gpu = self._prepare(left_images), self._prepare(right_images)
        return self._stereo_algo.compute_disparity_batch(left_gpu, right_gpu, out)

    @property
    def native(self) -> "cuda_depth.StereoMatching":
        return self._stereo_algo
