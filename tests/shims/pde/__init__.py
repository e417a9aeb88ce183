"""`import pde` -> marlpde_b200.pde_standin (see ../README.md)."""
from marlpde_b200.pde_standin import CartesianGrid, FieldCollection, ScalarField  # noqa: F401

__version__ = "0+standin"
