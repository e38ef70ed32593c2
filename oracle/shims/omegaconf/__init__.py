"""Stub (oracle only): `omegaconf.MISSING` as used by the reference's dataclass configs."""
MISSING = "???"
