"""CPU oracle for the m6anet MIL-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy (and in C under ``oracle/c``), the algorithm of the
reference path ``m6anet inference`` (reference ``m6anet/utils/inference_utils.py:14-104``,
``m6anet/model/model_blocks/blocks.py``, ``m6anet/model/model_blocks/pooling_blocks.py``).

Nothing in the product package ``m6anet_b200`` may import it.  The only permitted
importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``, where it is the checker / the timed CPU
baseline and never the thing shipped.

Parity status: PINNED.  ``tests/test_oracle.py`` checks the restatement against
 (a) the reference's own golden files for its bundled data (per-read probabilities,
     mod_ratio, site probabilities; copied as fixtures under ``tests/golden/bundled``),
 (b) outputs of the unmodified reference imported from ``/root/reference`` in the build
     container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``),
 (c) Random123 known-answer vectors for Philox4x32-10 and D. B. Thomas' MWC64X recurrence.
"""
from .philox import (philox4x32_10, sample_indices, sample_indices_many, sample_indices_mt19937,  # noqa: F401
                     block_layout, bag_indices, bag_indices_many, bag_indices_mt19937)
from .mil_oracle import (  # noqa: F401
    ReadEncoderParams,
    read_probabilities,
    read_probabilities_float64,
    noisy_or_site_probability,
    mod_ratio,
    mil_inference,
    closed_form_site_probability,
    pool_bags,
    mil_validate,
)
