"""CPU oracle for the DiffAssemble denoiser + sampling-loop hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffassemble_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline -- never as the product path.

PARITY UNPINNED.  The reference (IIT-PAVIS/DiffAssemble) ships no tests, golden
vectors or fixtures for this path, and the arithmetic of its attention layer
lives in ``torch_geometric.nn.TransformerConv`` which is neither vendored nor
version-pinned (``singularity/build/conda_env.yaml:12`` says just ``pyg``) and
is not installable in the build image (no network).  The reference package
itself cannot be imported here either (``pytorch_lightning``, ``timm``,
``pytorch3d`` ... are absent and ``puzzle_diff/model/backbones/__init__.py:1``
imports a module that does not exist).  This package therefore restates, line
by line and in plain torch on the CPU, the reference files named in each
docstring, plus the documented behaviour of the PyG / pytorch3d functions they
call.  It is anchored by algebraic known-answer tests (``tests/test_oracle_*``)
and by self-generated golden vectors (``tests/golden`` + the generating script).

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
from .assignment import greedy_cost_assignment_ref  # noqa: F401,E402
