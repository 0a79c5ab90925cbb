"""Helpers around mBGP model construction and training (MBGP/training_helpers.py): Phi constructors, optimiser
strategies, prediction summaries.  Host orchestration only -- every loss / gradient evaluation goes to the CUDA library."""
import abc
import copy
import dataclasses
import logging
from typing import Optional, Sequence, Tuple

import numpy as np

from .. import gpflow_shim as gpflow
from . import VBHelperFunctions
from .assigngp import AssignGP, BranchKernelParam, get_branching_point_kernel
from .data_generation import ColumnVector, GeneExpressionData, ToyBranchedData, Vector
from .FitBranchingModel import GetInitialConditionsAndPrior

InitialAndPriorPhi = Tuple[np.ndarray, np.ndarray]
_DEFAULT_NUM_PREDICTIONS = 100
_EARLY_BP = 1e-4
LOG = logging.getLogger("MMBGP training")


class PhiConstructor(abc.ABC):
    """Builds the initial and prior assignment probabilities."""

    @abc.abstractmethod
    def build(self) -> InitialAndPriorPhi:
        """Return (phi_initial [N,2], phi_prior [N,3])."""


def _with_trunk_column(phi_prior):
    return np.c_[np.zeros(phi_prior.shape[0])[:, None], phi_prior]


class SimplePhiConstructor(PhiConstructor):
    def __init__(self, gene_expression_data: GeneExpressionData, prior_confidence: float, allow_infs: bool = True) -> None:
        self._state = gene_expression_data.state
        self._prior_confidence = prior_confidence
        self._allow_infs = allow_infs

    def build(self) -> InitialAndPriorPhi:
        phi_initial, phi_prior = GetInitialConditionsAndPrior(self._state, self._prior_confidence, infPriorPhi=self._allow_infs)
        return phi_initial, _with_trunk_column(phi_prior)


def get_funky_phi(global_branching_state: Vector, uninformative_until: float, informative_prior_confidence: float) -> InitialAndPriorPhi:
    """Uninformative prior for the first `uninformative_until` fraction of cells (assumed uniformly spaced on [0, 1]),
    informative afterwards; initial phi as in SimplePhiConstructor (one draw from numpy's global RNG per branch cell)."""
    if not 0 <= uninformative_until <= 1:
        raise ValueError(f"Expected uninformative_until to be in the range [0, 1]. Instead got: {uninformative_until}.")
    if not 0 <= informative_prior_confidence <= 1:
        raise ValueError(f"Expected informative_prior_confidence to be in the range [0, 1]. Instead got: {uninformative_until}.")
    (N,) = global_branching_state.shape
    cut = int(uninformative_until * N)
    phi_initial = np.ones((N, 2)) * 0.5
    phi_prior = np.ones((N, 2)) * 0.5
    for i in range(N):
        if global_branching_state[i] in {2, 3}:
            col = int(global_branching_state[i] - 2)
            if i > cut:
                phi_prior[i, :] = 1 - informative_prior_confidence
                phi_prior[i, col] = informative_prior_confidence
            phi_initial[i, col] = 0.5 + (np.random.random() / 2.0)
            phi_initial[i, 1 - col] = 1 - phi_initial[i, col]
    assert np.allclose(phi_prior.sum(1), 1)
    assert np.allclose(phi_initial.sum(1), 1)
    return phi_initial, phi_prior


class FunkyPrior(PhiConstructor):
    def __init__(self, gene_expression_data: GeneExpressionData, uninformative_until: float, informative_prior_confidence: float,
                 allow_infs: bool = True) -> None:
        self._state = gene_expression_data.state
        self._uninformative_until = uninformative_until
        self._informative_prior_confidence = informative_prior_confidence
        self._allow_infs = allow_infs

    def build(self) -> InitialAndPriorPhi:
        phi_initial, phi_prior = get_funky_phi(self._state, self._uninformative_until, self._informative_prior_confidence)
        return phi_initial, _with_trunk_column(phi_prior)


class AssignGPOptimiser(abc.ABC):
    @abc.abstractmethod
    def train(self, model: AssignGP) -> AssignGP:
        """Optimise the model's trainable parameters and return the trained model."""


class ScipyOptimiser(AssignGPOptimiser):
    def __init__(self, method: str = "L-BFGS-B", maxiter: int = 100, **kwargs) -> None:
        self._method = method
        self._kwargs = kwargs
        if "maxiter" not in self._kwargs:
            self._kwargs["maxiter"] = maxiter
        else:
            assert maxiter == kwargs["maxiter"], f"Two different values of maxiter provided. Directly: {maxiter}, in kwargs: {kwargs['maxiter']}"

    def train(self, model: AssignGP) -> AssignGP:
        LOG.info(f"Starting training. Initial loss: {model.training_loss()}.")
        result = gpflow.optimizers.Scipy().minimize(model.training_loss, variables=model.trainable_variables, options=dict(self._kwargs),
                                                    method=self._method)
        LOG.debug(f"Optimisation result: {result}")
        LOG.info(f"Training complete. Final loss: {model.training_loss()}.")
        return model


def uniform_bp(model: AssignGP, bp: float = _EARLY_BP) -> AssignGP:
    model.kernel.Bv.assign(np.ones((model.BranchingPoints.shape[0],)) * bp)
    return model


def fixed_bps(model: AssignGP) -> AssignGP:
    gpflow.set_trainable(model.kernel.Bv, False)
    return model


def trainable_bps(model: AssignGP) -> AssignGP:
    gpflow.set_trainable(model.kernel.Bv, True)
    return model


def _clone(model: AssignGP) -> AssignGP:
    """gpflow.utilities.deepcopy: an independent copy of the parameters; the GPU engine handle is shared, not copied."""
    eng = model._engine
    model._engine = None
    try:
        new = copy.deepcopy(model)
    finally:
        model._engine = eng
    new._engine = eng
    return new


class ElvijsAmazingOptimiser(AssignGPOptimiser):
    """Train with all branching points pinned early, then restart from several uniform branching points and keep the best."""

    def __init__(self, base_optimiser: Optional[AssignGPOptimiser] = None, branching_points: Sequence[float] = (0.2, 0.4, 0.6, 0.8),
                 initial_bp: float = _EARLY_BP) -> None:
        self._base_optimiser = base_optimiser or ScipyOptimiser()
        self._bps = branching_points
        self._initial_bp = initial_bp

    def train(self, model: AssignGP) -> AssignGP:
        model = fixed_bps(uniform_bp(model, bp=self._initial_bp))
        model_with_reasonable_params = trainable_bps(self._base_optimiser.train(model))
        best_loss = model.training_loss()
        best_model = _clone(model)
        for bp in self._bps:
            new_model = uniform_bp(_clone(model_with_reasonable_params), bp)
            trained_model = self._base_optimiser.train(new_model)
            loss = trained_model.training_loss()
            if loss < best_loss:
                best_loss, best_model = loss, trained_model
        return best_model


class ElvijsRandomisingOptimiser(AssignGPOptimiser):
    """Train with early branching points, pick the best of `num_samples` uniformly sampled branching-point vectors, train again.
    The candidates are scored in batches by ONE call of the CUDA bound each (they differ only in their branching points)."""

    def __init__(self, base_optimiser: Optional[AssignGPOptimiser] = None, num_samples: int = 5_000, initial_bp: float = 1e-4,
                 batch: int = 256):
        self._num_samples = num_samples
        self._base_optimiser = base_optimiser or ScipyOptimiser()
        self._initial_bp = initial_bp
        self._batch = batch

    def train(self, model: AssignGP) -> AssignGP:
        num_outputs = model.kernel.Bv.shape[0]
        model.kernel.Bv.assign(np.ones((num_outputs,)) * self._initial_bp)
        gpflow.set_trainable(model.kernel.Bv, False)
        m = self._base_optimiser.train(model)
        gpflow.set_trainable(m.kernel.Bv, True)
        best_loss = m.training_loss()
        best_bps = m.kernel.Bv.numpy().copy()
        done = 0
        while done < self._num_samples:
            n = min(self._batch, self._num_samples - done)
            cand = np.random.uniform(size=(n, num_outputs))
            losses = m.training_losses_for_branching_points(cand)
            k = int(np.argmin(losses))
            if losses[k] < best_loss:
                best_loss, best_bps = float(losses[k]), cand[k].copy()
            done += n
        m.kernel.Bv.assign(best_bps)
        return self._base_optimiser.train(m)


class CompositeOptimiser(AssignGPOptimiser):
    def __init__(self, *optimisers: AssignGPOptimiser) -> None:
        self._optimisers = optimisers

    def train(self, model: AssignGP) -> AssignGP:
        for optimiser in self._optimisers:
            model = optimiser.train(model)
        return model


def construct_assigngp_model(gene_expression: GeneExpressionData, phi_constructor: PhiConstructor,
                             initial_branching_points: Sequence[float], kern: Optional[BranchKernelParam] = None) -> AssignGP:
    phi_initial, phi_prior = phi_constructor.build()
    x_expanded, indices, _ = VBHelperFunctions.GetFunctionIndexListGeneral(gene_expression.t)
    if not kern:
        kern = get_branching_point_kernel(branching_points=initial_branching_points)
    else:
        np.testing.assert_array_equal(kern.Bv.numpy(), np.array(initial_branching_points))
    return AssignGP(gene_expression.t, x_expanded, gene_expression.Y, indices, kern=kern, phiInitial=phi_initial, phiPrior=phi_prior,
                    multi=True)


def construct_and_fit_assigngp_model(gene_expression: GeneExpressionData, phi_constructor: PhiConstructor,
                                     initial_branching_points: Sequence[float], optimiser: AssignGPOptimiser,
                                     kern: Optional[BranchKernelParam] = None) -> AssignGP:
    model = construct_assigngp_model(gene_expression=gene_expression, phi_constructor=phi_constructor,
                                     initial_branching_points=initial_branching_points, kern=kern)
    return optimiser.train(model)


@dataclasses.dataclass
class GaussianPredictions:
    """Test points x and the predictive mean / variance of f (trunk), g and h (branches) at x, in this order."""
    x: ColumnVector
    y_mean: Tuple[ColumnVector, ColumnVector, ColumnVector]
    y_var: Tuple[ColumnVector, ColumnVector, ColumnVector]


@dataclasses.dataclass
class TrainingOutcome:
    model: AssignGP
    predictions: GaussianPredictions
    learned_branching_points: ColumnVector


def get_predictions(model: AssignGP) -> GaussianPredictions:
    grid = np.linspace(0, 1, _DEFAULT_NUM_PREDICTIONS)
    means, variances = [], []
    for f in range(1, 4):
        col = grid.reshape(_DEFAULT_NUM_PREDICTIONS, 1)
        mean, variance = model.predict_y(np.hstack((col, col * 0 + f)))
        means.append(np.asarray(mean))
        variances.append(np.asarray(variance))
    return GaussianPredictions(x=grid, y_mean=tuple(means), y_var=tuple(variances))


def get_training_outcome(trained_model: AssignGP) -> TrainingOutcome:
    return TrainingOutcome(model=trained_model, predictions=get_predictions(model=trained_model),
                           learned_branching_points=trained_model.kernel.Bv.numpy())


def get_assigngp_with_target_bps(bps: Sequence[float], lengthscale: float, noise_variance: float) -> AssignGP:
    data = ToyBranchedData(bps, N=100)
    m = construct_assigngp_model(gene_expression=data, phi_constructor=SimplePhiConstructor(data, prior_confidence=0.65),
                                 initial_branching_points=[0.5] * data.num_genes)
    m.kernel.Bv.assign(np.array(bps))
    m.kernel.kern.lengthscales.assign(lengthscale)
    m.likelihood.variance.assign(noise_variance)
    return m
