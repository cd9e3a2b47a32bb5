from .bcnn import BCNN, BilinearPooling  # noqa: F401
from .cbcnn import CBCNN, CompactBilinearPooling  # noqa: F401
