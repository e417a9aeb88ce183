"""`import h5py` -> marlpde_b200.hdf5lite (see ../README.md).  Surface used by the reference:
File(path, 'r').get(name) with .shape and indexing (test_regression.py:16-18, :52, :112, :139);
File(path, 'w') as a context manager with create_dataset(name, data=) and attrs.update(dict)
(Evolve_scenario.py:172-178)."""
from marlpde_b200.hdf5lite import Dataset, File  # noqa: F401

__version__ = "0+hdf5lite"
