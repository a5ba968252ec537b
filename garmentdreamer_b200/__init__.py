"""B200-native SDS hot path for GarmentDreamer: Gaussian-splat rasteriser + SD UNet step.

The product path is CUDA only (sm_100a shared libraries under ``garmentdreamer_b200/lib``);
there is no CPU fallback -- importing the bindings without the built libraries raises.
"""
__version__ = "0.1.0"
