"""Import alias: `import vlbert_b200` loads the package that lives in the directory `vl-bert_b200/`
(a hyphen cannot appear in a Python import statement)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vl-bert_b200")
_spec = importlib.util.spec_from_file_location("vlbert_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vlbert_b200"] = _mod
_spec.loader.exec_module(_mod)
