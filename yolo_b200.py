"""Import alias: `import yolo_b200` loads the package directory `yolo-compression-and-deployment-in-fpga_b200/`
(whose name is not a valid Python identifier) and registers it in sys.modules under this name, so that
`import yolo_b200.export`, `from yolo_b200 import lib` etc. work."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("yolo-compression-and-deployment-in-fpga_b200")
sys.modules[__name__] = _pkg
