"""stereo_depth_b200 -- B200-native (sm_100a) stereo-matching backend.

Drop-in for the `"cuda"` backend of dusanerdeljan/stereo-depth's DepthEstimationPipeline:
  * `cuda_depth`  : module with the reference extension's names (StereoMatchingConfiguration, StereoMatching)
  * `backend`     : StereoMatching ABC + CudaStereoMatchingBackend (process / process_batch)
  * `_native`     : ctypes binding of the C ABI in include/stereo_b200.h (libstereo_b200.so)
There is no CPU or PyTorch fallback: importing `_native` without the built library raises.
"""
__all__ = ["cuda_depth", "backend", "synthetic"]
