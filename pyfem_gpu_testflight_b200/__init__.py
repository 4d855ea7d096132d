"""pyfem_gpu_testflight_b200 -- B200-native finite-element assembly behind the pyfem API.

    import pyfem_gpu_testflight_b200 as pyfem

exposes the reference's names for the assembly path: Quadrature*/Basis* for quad4 and hex8,
LinearPoisson, LinearElasticity, Helmholtz, NonlinearPoisson2D, Assembler, ProblemCreator.
"""
from .assembler import Assembler
from .engine import DeviceMesh
from .fem import (BasisBase, BasisBilinear2D, BasisBlock3D, BasisBrick20Nodes, BasisTetrahedron10node,
                  BasisTriangle2D, QuadratureBase, QuadratureBilinear2D, QuadratureBlock3D,
                  QuadratureBrick333Point, QuadratureTetrahedron5Point, QuadratureTriangle2D)
from .mesh import ProblemCreator
from .models import Helmholtz, LinearElasticity, LinearPoisson, ModelBase, NonlinearPoisson2D

__all__ = [
    "Assembler", "DeviceMesh", "ProblemCreator", "ModelBase", "LinearPoisson", "LinearElasticity", "Helmholtz",
    "NonlinearPoisson2D", "QuadratureBase", "QuadratureBilinear2D", "QuadratureBlock3D", "BasisBase",
    "BasisBilinear2D", "BasisBlock3D",
]
