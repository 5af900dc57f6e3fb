"""Input side of the inference scripts: `TestingDataset(config)` yields (images, filenames) batches
(reference: lib_yolo/dataset_utils.py:188-219, TFRecords; here image files, see byolo.compat.ImageDataset)."""
from byolo.compat import ImageDataset as TestingDataset  # noqa: F401
