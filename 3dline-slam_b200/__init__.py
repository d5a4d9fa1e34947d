"""l3dpp-b200: B200-native Line3D++ matching / triangulation / scoring / affinity stage.

The directory name is not a valid Python identifier; load it with
``importlib.import_module("3dline-slam_b200")`` (tests/conftest.py and bench.py do).
"""
