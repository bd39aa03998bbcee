"""B200-native drop-in for the trace! -> segmentize! hot path of RayTracing.jl."""
from .mesh import DiscreteModelFromFile, GmshDiscreteModel, Mesh, UnstructuredDiscreteModel  # noqa: F401
from . import synth  # noqa: F401
