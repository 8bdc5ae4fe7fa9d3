"""Re-export of the seeded synthetic-field generator (spatpca_b200/synth.py) for the tests."""
from spatpca_b200.synth import case_names, make_case  # noqa: F401
