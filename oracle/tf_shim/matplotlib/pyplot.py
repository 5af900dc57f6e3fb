"""Import stub (plotting is out of scope)."""


def _unavailable(*a, **k):
    raise RuntimeError('matplotlib is not installed; plotting is out of scope')


imread = imshow = show = imsave = _unavailable
