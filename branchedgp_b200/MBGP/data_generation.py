"""Data containers and the toy generator used by the mBGP tests (MBGP/data_generation.py:33-308); plotting is out of scope."""
import dataclasses
from typing import Optional, Sequence

import numpy as np

Vector = np.ndarray        # shape (n,)
ColumnVector = np.ndarray  # shape (n, 1)
Matrix = np.ndarray        # shape (n, k)


class GeneExpressionDataValidationError(ValueError):
    """Raised when the data provided is internally inconsistent."""


@dataclasses.dataclass(frozen=True)
class GeneExpressionData:
    t: Vector                               # pseudotime, shape (N,)
    Y: Matrix                               # gene expression, shape (N, K)
    state: Vector                           # global state labels in {1, 2, 3}, shape (N,)
    gene_labels: Optional[Sequence[str]]    # K labels

    def __post_init__(self) -> None:
        self._validate()

    @property
    def num_genes(self) -> int:
        return self.Y.shape[1]

    def _validate(self) -> None:
        if len(self.t.shape) != 1:
            raise GeneExpressionDataValidationError(f"Expected pseudotime to have shape (N, ), instead got {self.t.shape}")
        if len(self.Y.shape) != 2:
            raise GeneExpressionDataValidationError(f"Expected gene expressions to have shape (N, K), instead got {self.Y.shape}")
        if self.gene_labels and len(self.gene_labels) != self.Y.shape[1]:
            raise GeneExpressionDataValidationError(f"Expected {self.Y.shape[1]} gene labels, instead got {len(self.gene_labels)}")
        if self.t.shape[0] != self.Y.shape[0] or self.state.shape != self.t.shape:
            raise GeneExpressionDataValidationError("t, Y and state must describe the same N cells")


@dataclasses.dataclass(frozen=True)
class BranchedData(GeneExpressionData):
    branching_points: Vector   # true branching point of every gene, shape (K,)

    def _validate(self) -> None:
        super()._validate()
        if len(self.branching_points) != self.Y.shape[1]:
            raise GeneExpressionDataValidationError("one branching point per gene is required")


class ToyBranchedData(BranchedData):
    """Noise-free quadratic branches +-(t^2 - b^2) after each gene's branching point; cells alternate between states 2 and 3."""

    def __init__(self, B=(0.1, 0.5, 0.8), N=20, gene_labels: Optional[Sequence[str]] = None):
        t = np.linspace(0, 1, N)
        Y = np.zeros((N, len(B)))
        state = np.ones(N, dtype=int)
        state[::2] = 2
        state[1::2] = 3
        for ib, b in enumerate(B):
            up = np.logical_and(t > b, state == 2)
            down = np.logical_and(t > b, state == 3)
            Y[up, ib] = t[up] ** 2 - b ** 2
            Y[down, ib] = -t[down] ** 2 + b ** 2
        super().__init__(t=t, Y=Y, branching_points=np.asarray(B, dtype=float), state=state, gene_labels=gene_labels)
