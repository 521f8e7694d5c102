"""CPU oracle for the MRefSR reference-alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``mrefsr_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline -- never as the product path.

What it is: a CPU restatement (torch fp32/fp64 tensor ops + a plain-C DCN in
``dcn_ref.c``) of the reference algorithms

* correspondence matcher      basicsr/archs/ref_map_util.py:4-86
* idx -> flow -> pre-offsets   basicsr/archs/corres_generation_arch.py:30-105,
                               basicsr/archs/arch_util.py:386-410
* DynAgg offset/mask glue      basicsr/archs/ref_mrapa_restoration_arch.py:45-76
* modulated deformable conv    basicsr/ops/dcn/src/deform_conv_cuda_kernel.cu:468-767,
                               basicsr/ops/dcn/src/deform_conv_cuda.cpp:490-685
* multi-reference attention    basicsr/archs/ref_mrapa_restoration_arch.py:321-335

Pinning: the reference ships NO golden vectors / known-answer tests for this
path, so the oracle is pinned against outputs of the
reference code itself, run in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference unmodified) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
every oracle function against those fixtures.
"""
from .matcher import feature_match_index_oracle, sample_patches_oracle, similarity_volume  # noqa: F401
from .correspondence import index_to_flow_oracle, pre_offsets_oracle  # noqa: F401
from .dcn import (  # noqa: F401
    modulated_deform_conv_oracle,
    modulated_deform_conv_backward_oracle,
    dynagg_offsets_oracle,
)
from .fusion import mrapa_attention_oracle  # noqa: F401
