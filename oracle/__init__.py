"""CPU oracle for the Restormer / DnCNN forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it is a
CPU restatement of the reference's algorithm (leducthanhig/image-restoration-models,
``src/restormer/restormer.py`` and ``src/dncnn/models/network_dncnn.py``) that is
imported solely by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The shipped
package (``image_restoration_models_b200``) never imports it and has no CPU
fallback.

Parity status: PINNED.  The reference publishes no golden vectors of its own
(SURVEY.md §8c), so the oracle is pinned against outputs of the unmodified
reference modules imported in the build container from ``/root/reference/src``:
``oracle/make_golden.py`` writes those outputs to ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks the restatement against them.
"""
from .synth import (synth_tensor, synth_image, restormer_schema, dncnn_schema,
                    synth_state_dict, RESTORMER_TASKS)
from .restormer_ref import restormer_forward, transformer_block, layer_norm, attention, feed_forward
from .dncnn_ref import dncnn_forward

__all__ = [
    "synth_tensor", "synth_image", "restormer_schema", "dncnn_schema", "synth_state_dict",
    "RESTORMER_TASKS", "restormer_forward", "transformer_block", "layer_norm", "attention",
    "feed_forward", "dncnn_forward",
]
