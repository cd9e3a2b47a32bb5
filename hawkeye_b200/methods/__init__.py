from .bcnn import BCNN, BilinearPooling  # noqa: F401
