"""Importable alias of the package directory
`multimodal-hierarchical-cnn-for-sun-salutation-pose-classification_b200/` (its name contains hyphens, so it
cannot be imported directly). `import qtcnn_b200.capi`, `qtcnn_b200.models`, ... resolve to the modules there.
"""
import os as _os

PACKAGE_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                            "multimodal-hierarchical-cnn-for-sun-salutation-pose-classification_b200")
__path__.append(PACKAGE_DIR)
__version__ = "0.1.0"
