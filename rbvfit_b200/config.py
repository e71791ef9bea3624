"""Host-side mirror of the part of rbvfit's FitConfiguration that the hot path consumes
(reference: src/rbvfit/core/fit_configuration.py:24-351, 380-391, 472-494).

``GpuVoigtModel`` accepts either this class or the reference's own ``FitConfiguration`` -- it only reads
``config.systems[*].redshift``, ``.ion_groups[*].transitions/.components/.ion_name``,
``config.instrumental_params`` and calls ``config.validate()`` (voigt_model.py:367-374, 391-401, 428-437).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List

from . import lines


@dataclass
class IonGroup:
    ion_name: str
    transitions: List[float]
    components: int
    redshift: float
    validate_ion: bool = True

    def __post_init__(self):
        if self.components <= 0:
            raise ValueError(f"Number of components must be positive, got {self.components}")
        if not self.transitions:
            raise ValueError(f"No transitions provided for ion {self.ion_name}")
        snapped, detected = [], []
        for wave in self.transitions:      # snap to database wavelengths (fit_configuration.py:104-124)
            info = lines.rb_setline(wave, "closest")
            exact = info["wave"][0]
            ion = lines.ion_of(info["name"][0])
            if self.validate_ion and ion != self.ion_name:
                raise ValueError(f"Invalid transition {wave}Å: Transition {wave}Å (corrected to {exact}Å) "
                                 f"belongs to {ion}, not {self.ion_name}")
            snapped.append(exact)
            detected.append(ion)
        self.transitions = snapped

    def get_parameter_count(self) -> int:
        return 3 * self.components


@dataclass
class AbsorptionSystem:
    redshift: float
    ion_groups: List[IonGroup] = field(default_factory=list)

    def add_ion(self, ion_name, transitions, components, merge=False, validate_ion=True):
        for group in self.ion_groups:
            if group.ion_name == ion_name:
                if not merge:
                    raise ValueError(f"Ion {ion_name} already exists in system at z={self.redshift}. "
                                     f"Use merge=True to add transitions to existing ion.")
                if group.components != components:
                    raise ValueError(f"Cannot merge ion {ion_name}: component mismatch "
                                     f"(existing: {group.components}, new: {components})")
                extra = IonGroup(ion_name, list(transitions), components, self.redshift, validate_ion)
                for w in extra.transitions:
                    if not any(abs(w - e) < 1e-3 for e in group.transitions):
                        group.transitions.append(w)
                return
        self.ion_groups.append(IonGroup(ion_name, list(transitions), components, self.redshift, validate_ion))

    def get_parameter_count(self) -> int:
        return sum(g.get_parameter_count() for g in self.ion_groups)


class FitConfiguration:
    """Systems -> ion groups -> transitions x components; parameters tied within an ion group."""

    def __init__(self, FWHM=None, grating=None, life_position=None, cen_wave=None):
        self.systems: List[AbsorptionSystem] = []
        self._validated = False
        self.instrumental_params: Dict[str, Any] = {}
        for key, val in (("FWHM", FWHM), ("grating", grating), ("life_position", life_position),
                         ("cen_wave", cen_wave)):
            if val is not None:
                self.instrumental_params[key] = val

    def add_system(self, z, ion="auto", transitions=None, components=1, merge=False, validate_ion=True):
        if transitions is None:
            raise ValueError("transitions list cannot be None")
        system = None
        for s in self.systems:                      # fit_configuration.py:380-391
            if abs(s.redshift - z) < 1e-6:
                system = s
                break
        if system is None:
            system = AbsorptionSystem(z)
            self.systems.append(system)
        if ion == "auto":
            ion = lines.ion_of(lines.rb_setline(transitions[0], "closest")["name"][0])
        system.add_ion(ion, transitions, components, merge, validate_ion)
        self._validated = False

    def validate(self) -> None:
        if self._validated:
            return
        if not self.systems:
            raise ValueError("No absorption systems defined")
        for system in self.systems:
            if not system.ion_groups:
                raise ValueError(f"System at z={system.redshift} has no ions defined")
        self._validated = True

    def get_parameter_structure(self) -> Dict[str, Any]:
        total = sum(s.get_parameter_count() for s in self.systems)
        return {"total_parameters": total,
                "systems": [{"redshift": s.redshift,
                             "ion_groups": [{"ion": g.ion_name, "transitions": list(g.transitions),
                                             "components": g.components} for g in s.ion_groups]}
                            for s in self.systems]}

    def __repr__(self):
        n = sum(len(s.ion_groups) for s in self.systems)
        return f"FitConfiguration({len(self.systems)} systems, {n} ion groups)"
