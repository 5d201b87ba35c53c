"""CPU oracle for the DiffAssemble denoiser + sampling-loop hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffassemble_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline -- never as the product path.

PARITY: PINNED AGAINST THE REFERENCE'S OWN CODE, EXCEPT TWO UN-VENDORED THIRD-PARTY
FUNCTIONS.  The reference (IIT-PAVIS/DiffAssemble) ships no tests or golden vectors,
but its Python can be executed in the build container once the absent packages
(``pytorch_lightning``, ``timm``, ``torchmetrics``, ``kornia`` ...) are replaced by inert
placeholders: ``tests/golden/make_reference_golden.py`` imports
``/root/reference/puzzle_diff/model`` and drives the real ``GNN_Diffusion`` (2-D and
3-D), ``Eff_GAT``, ``Eff_GAT_3d``, ``Transformer_GNN``, ``Exophormer_GNN``, the samplers,
``p_losses`` + autograd, ``greedy_cost_assignment`` and the topology generators on seeded
inputs; ``tests/test_oracle_pinned.py`` holds this package to those vectors (2e-6
relative, exact for indices) and ``tests/test_gpu_reference_vectors.py`` holds the CUDA
path to them.  STILL UNPINNED: ``torch_geometric.nn.TransformerConv`` and
``pytorch3d.transforms.matrix_to_quaternion / quaternion_to_matrix`` -- neither vendored
nor version-pinned by the reference (``singularity/build/conda_env.yaml:12`` says just
``pyg``; pytorch3d is in no env file) and not installable here -- are restated from their
documented behaviour in ``transformer_conv.py`` / ``so3.py``, supplied to the reference
run by the same restatement, and anchored only by the algebraic known-answer tests in
``tests/test_oracle_kats.py`` (dense == scaled-dot-product attention, uniform attention
at W_q = 0, single in-edge, isolated node, duplicate edge, the 1e-16 denominator ...).

All ``path:line`` citations are relative to the reference checkout.
"""

from .transformer_conv import TransformerConvRef, segment_softmax  # noqa: F401
from .gnn import TransformerGNNRef, ExophormerGNNRef  # noqa: F401
from .eff_gat import EffGATRef, EffGAT3dRef  # noqa: F401
from .diffusion import (  # noqa: F401
    ModelMeanType,
    ModelScheduler,
    GNNDiffusionRef,
    GNNDiffusion3dRef,
    linear_beta_schedule,
    cosine_beta_schedule,
    cosine_discrete_beta_schedule,
    extract,
)
from .topology import (  # noqa: F401
    dense_edge_index,
    generate_random_regular_graph,
    generate_random_expander,
    batch_graphs,
)
from .assignment import greedy_cost_assignment_ref, puzzle_accuracy_ref  # noqa: F401,E402
from .pointnet import PointNetRef  # noqa: F401,E402
