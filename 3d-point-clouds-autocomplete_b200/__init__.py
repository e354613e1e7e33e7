"""B200-native (sm_100a) point-set hot path of HyperPocket (gmum/3d-point-clouds-autocomplete).

The directory name is not a Python identifier; import it with
``importlib.import_module("3d-point-clouds-autocomplete_b200")`` or through the ``hp_b200``
shim at the repository root.  ``dropin/`` mirrors the reference's import paths
(``losses.champfer_loss``, ``utils.pytorch_structural_losses.*``, ``model.target_network``,
``utils.metrics``) for use on ``sys.path`` ahead of the reference tree.
"""
from . import _native  # noqa: F401
from .chamfer import (ChamferLoss, NNDistance, NNDistanceFunction, NNDistanceGrad, batch_pairwise_dist,  # noqa: F401
                      chamfer_backward, chamfer_forward, chamfer_step, chamfer_step_supported, nn_distance)

from .emd import (ApproxMatch, MatchCost, MatchCostFunction, MatchCostGrad, approx_match, emd_cost_pairs,  # noqa: F401
                  match_cost)

from . import target_network  # noqa: F401
from .target_network import (TargetNetwork, generate_points, generate_points_batched, reconstruct_batch,  # noqa: F401
                             target_network_backward, target_network_forward, target_network_num_weights,
                             target_network_set_mode)
from . import graphs  # noqa: F401
from .graphs import (ChamferHostPipeline, ChamferStepGraph, FullModelStepGraph, HotPathStepGraph,  # noqa: F401
                     TargetNetworkStepGraph)
from . import hyper_network  # noqa: F401
from .hyper_network import FusedHyperNetworkHead, fuse_hypernetwork_head  # noqa: F401

from . import metrics  # noqa: F401
from .metrics import compute_all_metrics, pairwise_cd, pairwise_emd  # noqa: F401

from . import evaluation  # noqa: F401

__version__ = "0.1.0"
