"""Stub (oracle only): mjtObj enum values (MuJoCo 3.x numbering) named by reference `utilities.py:25-39`."""


class mjtObj:
    mjOBJ_BODY = 1
    mjOBJ_JOINT = 3
    mjOBJ_SITE = 6
    mjOBJ_CAMERA = 7
    mjOBJ_SENSOR = 18
    mjOBJ_NUMERIC = 19
    mjOBJ_KEY = 23
