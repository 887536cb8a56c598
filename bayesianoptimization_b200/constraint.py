"""ConstraintModel on the device: the reference's class (R/bayes_opt/constraint.py:16-263, imported, not
restated) with its private sklearn GPs (:72-81) replaced by B200GaussianProcessRegressor objects of the
same configuration.  ``fit`` / ``approx`` / ``allowed`` / ``eval`` are inherited; ``predict`` is inherited
too (one device predict per constraint GP) and is only used stand-alone - inside an acquisition closure
the product of the probabilities is evaluated by the fused kernel (fused.FusedAcquisition)."""
from __future__ import annotations

from bayes_opt.constraint import ConstraintModel as _RefConstraintModel

from .gpr import to_b200_gp


class ConstraintModel(_RefConstraintModel):
    def __init__(self, fun, lb, ub, transform=None, random_state=None, device=0, devices=None):
        super().__init__(fun, lb, ub, transform=transform, random_state=random_state)
        self._model = [to_b200_gp(g, device, devices) for g in self._model]
