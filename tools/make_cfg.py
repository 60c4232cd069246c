"""Writes the packaged group-table yaml files from ms_hgnn.morphology.GROUPS."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "morphsym-hgnn_b200"))
import yaml
from ms_hgnn.morphology import CFG_DIR, GROUPS

class _NoAlias(yaml.SafeDumper):
    def ignore_aliases(self, data):
        return True


os.makedirs(CFG_DIR, exist_ok=True)
for name, g in GROUPS.items():
    with open(os.path.join(CFG_DIR, name + ".yaml"), "w") as f:
        yaml.dump(g, f, Dumper=_NoAlias, default_flow_style=None, sort_keys=False)
    print("wrote", name)
