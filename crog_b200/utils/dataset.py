"""Host-side data helpers that sit either side of the hot path (SURVEY.md §8 "next" rows), under the reference's module name
``utils/dataset.py``: ``tokenize`` (utils/dataset.py:57-98) lives in ``simple_tokenizer``; the letterbox matrices and the device
pre-processing (utils/dataset.py:825-866) live in ``warp``.  Nothing else of the reference's dataset code (LMDB readers,
augmentation, target-map generation) is built."""
from .simple_tokenizer import tokenize  # noqa: F401
from .warp import get_transform_mat, preprocess_images  # noqa: F401
