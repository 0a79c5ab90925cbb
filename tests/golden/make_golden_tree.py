"""Generates tests/golden/tree_index_list.npz from the REFERENCE's BinaryBranchingTree (numpy only, importable here):
GetFunctionIndexList / GetFunctionDomains of the one-branch-point tree for a seeded input, with the global RNG seeded the same way
the test seeds it.  Run from the repo root: python tests/golden/make_golden_tree.py"""
import importlib.util
import os

import numpy as np

np.NAN = np.nan   # the reference predates numpy 2
spec = importlib.util.spec_from_file_location("ref_bt", "/root/reference/BranchedGP/BranchingTree.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    x = np.random.RandomState(0).rand(40) * 0.99 + 0.005
    t = ref.BinaryBranchingTree(0, 1)
    t.add(None, 1, 0.4)
    np.random.seed(3)
    Xnew, idx, Xtrue = t.GetFunctionIndexList(x, fReturnXtrue=True)
    flat = np.array([i for ix in idx for i in ix])
    lens = np.array([len(ix) for ix in idx])
    fm, fmb = t.GetFunctionBranchTensor()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_index_list.npz"), x=x, Xnew=Xnew, idx_flat=flat,
                        idx_len=lens, Xtrue=Xtrue, domains=t.GetFunctionDomains(), fm=fm)
    print(Xnew.shape, lens.sum())


if __name__ == "__main__":
    main()
