"""Pre-training side of the hot path: the discrete low-resolution simulation of the MultiRes trainers on the GPU."""
from .discrete_downsampling import (SimulateDiscreteLowResolutionTransform,  # noqa: F401
                                    augment_discrete_linear_downsampling_scipy, resize_edge)
