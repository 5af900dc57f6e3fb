"""Drop-in for the reference's `lib_yolo` package, hot path only (model classes, prior tables, containers);
training, augmentation and TFRecord code of the reference are out of scope."""
