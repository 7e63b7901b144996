"""remo3d_b200: B200-native forward solve of ReMo3D.  `from remo3d_b200 import Model`."""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "Model":
        from .remo3d import Model

        return Model
    raise AttributeError(name)
