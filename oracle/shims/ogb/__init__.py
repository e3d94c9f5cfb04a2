"""Placeholder for `ogb` (datasets/evaluators only; never on the hot path). oracle/ test infrastructure."""
