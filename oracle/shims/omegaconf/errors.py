"""Stub (oracle only)."""


class MissingMandatoryValue(Exception):
    pass


class ConfigAttributeError(Exception):
    pass
