from .gpu_pipeline import (GpuTrainPipeline, image_prep, mask_flip_pad,  # noqa: F401
                           polygons_to_bitmaps, resize_bilinear_u8, resize_nearest_u8)
from .synthetic import make_inputs  # noqa: F401
from .bonai import BONAI, DATASETS, build_dataset  # noqa: F401
