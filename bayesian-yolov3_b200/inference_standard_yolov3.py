"""Same call surface as the reference's inference_standard_yolov3.py; the implementation is shared (byolo/inference_common.py)."""
from byolo import inference_common as _common

globals().update(_common.surface('standard'))

if __name__ == '__main__':
    _common.run_as_script(main)  # noqa: F821
