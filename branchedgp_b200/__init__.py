"""B200-native evaluator for the BranchedGP hot path (AssignGP variational bound + gradient).

The public modules mirror the reference's names: ``FitBranchingModel``, ``VBHelperFunctions``,
``assigngp_dense``, ``assigngp_denseSparse``, ``pZ_construction_singleBP``, ``branch_kernParamGPflow``,
``BranchingTree`` and ``MBGP``.  All bound / gradient arithmetic runs in ``libbgp_b200.so`` (CUDA, sm_100a).
"""
__version__ = "0.1.0"
