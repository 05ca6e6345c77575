"""Source root of the B200-native scattering path (`tebscat` python package + csrc/).

The directory name is not a python identifier; add it to sys.path and
``import tebscat`` (tests/conftest.py, bench.py and __graft_entry__.py do)."""
