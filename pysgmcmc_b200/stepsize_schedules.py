"""Stepsize schedules -- same interface as pysgmcmc/stepsize_schedules.py:4-91.

Host-side scalars: a sampler calls ``next(schedule)`` before and
``schedule.update(params, cost)`` after every step
(pysgmcmc/samplers/base_classes.py:195-197,306,443).
"""
from abc import ABCMeta, abstractmethod


class StepsizeSchedule(object, metaclass=ABCMeta):
    """Base class of all stepsize schedules (stepsize_schedules.py:4-34)."""

    def __init__(self, initial_value):
        self.initial_value = initial_value

    @abstractmethod
    def __next__(self):
        """Return the stepsize to use for the next sampler step."""

    def __iter__(self):
        return self

    @abstractmethod
    def update(self, *args, **kwargs):
        """Feed information about the last step (sample, cost, ...) back into the schedule."""


class ConstantStepsizeSchedule(StepsizeSchedule):
    """Keeps the stepsize at its initial value (stepsize_schedules.py:37-91).

    >>> schedule = ConstantStepsizeSchedule(0.01)
    >>> next(schedule)
    0.01
    >>> str(schedule)
    'ConstantStepsizeSchedule(stepsize=0.01)'
    """

    def __next__(self):
        return self.initial_value

    def __str__(self):
        return "ConstantStepsizeSchedule(stepsize={})".format(self.initial_value)

    def update(self, *args, **kwargs):
        pass
