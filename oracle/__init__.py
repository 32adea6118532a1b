"""CPU oracle for the Enhanced-UNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it,
and there only as the checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``enhanced_unet_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against the reference *itself*: ``oracle/make_golden.py`` imports ``/root/reference`` in the
build container, runs it on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks the restatement in this package against those fixtures (and
against the live reference when ``/root/reference`` is present).
"""
from .unet_oracle import (  # noqa: F401
    PARAM_SPECS,
    BUFFER_SPECS,
    make_state_dict,
    make_input,
    make_target,
    unet_forward,
    combined_loss,
    batch_loss,
    fusion_forward,
)
from .metrics_oracle import (  # noqa: F401
    confusion_counts,
    calculate_iou,
    calculate_dice,
    calculate_semantic_metrics,
    metrics_from_counts,
    convert_probs_to_mask,
    calculate_instance_metrics,
    make_instance_case,
    INSTANCE_CASES,
)
