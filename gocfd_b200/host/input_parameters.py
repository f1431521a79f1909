"""YAML input parameters for the 2D solver (host side).

Same schema and defaults as the reference.  The reference unmarshals through ghodss/yaml,
which goes YAML -> JSON -> encoding/json: keys are matched to the Go *field names*
case-insensitively (the `yaml:"..."` tags are ignored), unknown keys are dropped and the last
duplicate key wins (the shipped run scripts rely on that by appending overrides).

Reference: InputParameters/InputParameters.go:11-32, cmd/2D.go:63-103 (defaults Gamma=1.4,
Minf=0.1).
"""
from dataclasses import dataclass, field, fields
from typing import List

import yaml


@dataclass
class InputParameters2D:
    Title: str = ""
    CFL: float = 0.0
    FluxType: str = ""
    InitType: str = ""
    PolynomialOrder: int = 0
    FinalTime: float = 0.0
    Minf: float = 0.0
    Gamma: float = 0.0
    Alpha: float = 0.0
    BCs: dict = field(default_factory=dict)
    LocalTimeStepping: bool = False
    MaxIterations: int = 0
    ImplicitSolver: bool = False
    Limiter: str = ""
    Kappa: float = 0.0
    PlotFields: List[str] = field(default_factory=list)

    def parse(self, data):
        """Overlay YAML text onto this struct (ip.Parse)."""
        class _LastWins(yaml.SafeLoader):
            pass

        def construct_mapping(loader, node, deep=False):
            loader.flatten_mapping(node)
            return {loader.construct_object(k, deep=deep): loader.construct_object(v, deep=deep)
                    for k, v in node.value}

        _LastWins.add_constructor(yaml.resolver.BaseResolver.DEFAULT_MAPPING_TAG, construct_mapping)
        doc = yaml.load(data, Loader=_LastWins) or {}
        by_lower = {f.name.lower(): f for f in fields(self)}
        for key, val in doc.items():
            f = by_lower.get(str(key).lower())
            if f is None or val is None:
                continue
            if f.type is float:
                val = float(val)
            elif f.type is int:
                val = int(val)
            elif f.type is bool:
                val = bool(val)
            elif f.type is str:
                val = str(val)
            setattr(self, f.name, val)
        return self


def process_input(ic_file=None, n_order=-1):
    """cmd/2D.go processInput: defaults, then the YAML file if given."""
    ip = InputParameters2D(Gamma=1.4, Minf=0.1)
    if ic_file:
        with open(ic_file, "r") as f:
            ip.parse(f.read())
    else:
        ip.PolynomialOrder = n_order
        ip.FluxType = "Lax"
        ip.InitType = "Freestream"
    return ip
