from .gpu_pipeline import GpuTrainPipeline, image_prep, mask_flip_pad  # noqa: F401
from .synthetic import make_inputs  # noqa: F401
