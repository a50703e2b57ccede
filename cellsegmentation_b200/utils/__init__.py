from .image_processing import (generate_masks, heatmap, heatmap_arrays, hsv_refine_batch,  # noqa: F401
                               preprocess_masks, remove_small_regions)
