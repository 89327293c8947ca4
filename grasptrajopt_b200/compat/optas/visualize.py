"""Placeholder for ``optas.visualize`` (reference: ``optas/visualize.py``, VTK viewer).

Rendering is outside the hot path.  The module imports without ``vtk`` so that
``from optas.visualize import Visualizer`` in the example scripts succeeds; constructing the
viewer without VTK raises a clear error (the examples only do so under ``--vis``)."""


class Visualizer:
    def __init__(self, *args, **kwargs):
        try:
            import vtk  # noqa: F401
        except ImportError as exc:
            raise ImportError("optas.visualize.Visualizer needs the 'vtk' package; visualisation is not part of the B200 hot path") from exc
        raise NotImplementedError("VTK visualisation is out of scope of the B200 build; use the reference's optas/visualize.py")
