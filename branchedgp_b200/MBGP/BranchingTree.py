from ..BranchingTree import BinaryBranchingTree  # noqa: F401  (MBGP/BranchingTree.py is a copy of BranchingTree.py in the reference)
