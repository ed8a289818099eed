"""Exception types with the reference's names and hierarchy (opendrift/errors.py:5-15)."""


class WrongMode(Exception):
    pass


class NotCoveredError(Exception):
    pass


class OutsideSpatialCoverageError(NotCoveredError):
    pass


class OutsideTemporalCoverageError(NotCoveredError):
    pass


class VariableNotCoveredError(NotCoveredError):
    pass
