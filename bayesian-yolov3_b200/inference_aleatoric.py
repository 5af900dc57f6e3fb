"""Same call surface as the reference's inference_aleatoric.py; the implementation is shared (byolo/inference_common.py)."""
from byolo import inference_common as _common

globals().update(_common.surface('aleatoric'))

if __name__ == '__main__':
    _common.run_as_script(main)  # noqa: F821
