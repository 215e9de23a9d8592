"""See matplotlib/__init__.py: imported by the reference driver, never called on the inference path."""


def __getattr__(name):
    raise AttributeError('matplotlib.pyplot.%s: matplotlib is not installed; this stand-in only satisfies the import' % name)
