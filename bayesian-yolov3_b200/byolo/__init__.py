"""byolo: B200-native (sm_100a) detection hot path of flkraus/bayesian-yolov3 behind a C ABI (include/byolo.h)."""
from . import priors, weights  # noqa: F401  (pure numpy, importable without a GPU)

__all__ = ['priors', 'weights', 'Engine', 'nms', 'nms_per_class', 'conv_layer']


def __getattr__(name):
    if name in ('Engine', 'nms', 'nms_per_class', 'conv_layer'):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
