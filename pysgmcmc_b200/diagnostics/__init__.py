from . import objective_functions  # noqa: F401
