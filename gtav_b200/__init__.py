"""Importable alias of the `ai-generated-gtav_b200/` package directory (whose name is not a valid
Python identifier): `import gtav_b200.model.dit` loads ai-generated-gtav_b200/model/dit.py."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ai-generated-gtav_b200"))
__doc__ = open(_os.path.join(__path__[-1], "__init__.py")).read().split('"""')[1]
