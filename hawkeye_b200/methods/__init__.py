from .bcnn import BCNN, BilinearPooling  # noqa: F401
from .cbcnn import CBCNN, CompactBilinearPooling  # noqa: F401
from .mpn import MPN, MPNCOV  # noqa: F401
from .peer_learning import PeerLearningNet  # noqa: F401
from .cin import CIN, ChannelInteractionModule, CINClassifier  # noqa: F401
from .osme import OSMENet, OSME, OSME_block  # noqa: F401
