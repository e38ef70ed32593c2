from .lqr import LinearQuadraticRegulator, LinearQuadraticRegulatorConfig  # noqa: F401
