"""phybin_b200 -- B200-native all-to-all Robinson-Foulds distance matrices (PhyBin's hashRF /
naiveDistMatrix path).  See DESIGN.md."""
from .host import Forest, RANDOM_TOPOLOGY, NNI_FAMILY  # noqa: F401
