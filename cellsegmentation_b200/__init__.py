"""cellsegmentation_b200 — B200-native (sm_100a) Stage-2 MIL hot path + Stage-3 HSV refinement of
Newiz430/CellSegmentation behind the reference's own Python surfaces.

    from cellsegmentation_b200.model import nets            # model/__init__.py
    from cellsegmentation_b200.dataset import LystoDataset, LystoTestset, get_tiles
    from cellsegmentation_b200.inference import inference_tiles, sample, rank
    from cellsegmentation_b200.train import train_tile
    from cellsegmentation_b200.evaluate import evaluate_tile
    from cellsegmentation_b200.utils import heatmap, generate_masks, preprocess_masks

Every kernel lives in csrc/ and is reached through the C ABI in include/cellseg_b200.h.
"""
__version__ = "0.1.0"
