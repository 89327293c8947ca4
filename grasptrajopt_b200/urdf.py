"""Minimal URDF reader (xml.etree only).

The reference parses URDFs with the third-party ``urdf_parser_py`` package
(``optas/models.py:15,286-293``), which is not available offline.  This module
reads the subset of URDF the hot path needs -- links (visual origin + mesh
file), joints (type, parent/child, origin, axis, limits) -- and keeps the
conventions the reference relies on:

* joints/links are kept in document order; the actuated-joint index of a joint
  is its position among the non-fixed joints (``optas/models.py:349-354,661-667``),
* a joint without ``<axis>`` defaults to (1,0,0) (``optas/models.py:653-659``),
* a ``<limit>`` element without ``lower``/``upper`` yields 0.0 for them
  (urdf_parser_py behaviour noted in the reference ``README.md:121``); a joint
  without any ``<limit>`` yields -1e9/+1e9 (``optas/models.py:438-465``).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional


def _floats(text: Optional[str], n: int, default: float = 0.0) -> List[float]:
    if text is None:
        return [default] * n
    vals = [float(v) for v in text.replace(",", " ").split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return vals


@dataclass
class Pose:
    xyz: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    rpy: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])


@dataclass
class JointLimit:
    lower: float = 0.0
    upper: float = 0.0
    velocity: float = 0.0
    effort: float = 0.0


@dataclass
class Mesh:
    filename: str
    scale: Optional[List[float]] = None


@dataclass
class Visual:
    origin: Optional[Pose]
    geometry: Optional[Mesh]


@dataclass
class Link:
    name: str
    visual: Optional[Visual] = None


@dataclass
class Joint:
    name: str
    type: str
    parent: str
    child: str
    origin: Optional[Pose] = None
    axis: Optional[List[float]] = None
    limit: Optional[JointLimit] = None
    mimic: Optional[str] = None

    # urdf_parser_py spells it ``joint_type`` in the constructor and ``type`` as attribute.
    @property
    def joint_type(self) -> str:
        return self.type


def _parse_pose(elem) -> Optional[Pose]:
    if elem is None:
        return None
    return Pose(xyz=_floats(elem.get("xyz"), 3), rpy=_floats(elem.get("rpy"), 3))


class URDF:
    """Kinematic tree of a robot description."""

    def __init__(self, name: str = "robot"):
        self.name = name
        self.links: List[Link] = []
        self.joints: List[Joint] = []
        self.link_map: Dict[str, Link] = {}
        self.joint_map: Dict[str, Joint] = {}
        self.parent_map: Dict[str, tuple] = {}  # child link -> (joint name, parent link)

    # -- construction -------------------------------------------------------------------
    @classmethod
    def from_xml_string(cls, text: str) -> "URDF":
        if isinstance(text, bytes):
            text = text.decode("utf-8")
        root = ET.fromstring(text.encode("utf-8"))
        if root.tag != "robot":
            raise ValueError("URDF root element must be <robot>")
        model = cls(root.get("name", "robot"))
        for le in root.findall("link"):
            vis = None
            ve = le.find("visual")
            if ve is not None:
                mesh = None
                ge = ve.find("geometry")
                if ge is not None and ge.find("mesh") is not None:
                    me = ge.find("mesh")
                    scale = _floats(me.get("scale"), 3) if me.get("scale") else None
                    mesh = Mesh(me.get("filename"), scale)
                vis = Visual(_parse_pose(ve.find("origin")), mesh)
            model.add_link(Link(le.get("name"), vis))
        for je in root.findall("joint"):
            ae = je.find("axis")
            lime = je.find("limit")
            lim = None
            if lime is not None:
                lim = JointLimit(
                    lower=float(lime.get("lower", 0.0)),
                    upper=float(lime.get("upper", 0.0)),
                    velocity=float(lime.get("velocity", 0.0)),
                    effort=float(lime.get("effort", 0.0)),
                )
            mim = je.find("mimic")
            model.add_joint(
                Joint(
                    name=je.get("name"),
                    type=je.get("type"),
                    parent=je.find("parent").get("link"),
                    child=je.find("child").get("link"),
                    origin=_parse_pose(je.find("origin")),
                    axis=_floats(ae.get("xyz"), 3) if ae is not None else None,
                    limit=lim,
                    mimic=mim.get("joint") if mim is not None else None,
                )
            )
        return model

    @classmethod
    def from_xml_file(cls, filename: str) -> "URDF":
        with open(os.fspath(filename), "r", encoding="utf-8") as fh:
            return cls.from_xml_string(fh.read())

    def add_link(self, link: Link) -> None:
        self.links.append(link)
        self.link_map[link.name] = link

    def add_joint(self, joint: Joint) -> None:
        self.joints.append(joint)
        self.joint_map[joint.name] = joint
        self.parent_map[joint.child] = (joint.name, joint.parent)

    # -- queries ------------------------------------------------------------------------
    def get_root(self) -> str:
        roots = [l.name for l in self.links if l.name not in self.parent_map]
        if len(roots) != 1:
            raise ValueError(f"URDF must have exactly one root link, found {roots}")
        return roots[0]

    def get_chain(self, root: str, tip: str, joints: bool = True, links: bool = True) -> List[str]:
        """Names on the path root -> tip (same flags as urdf_parser_py)."""
        chain: List[str] = []
        if links:
            chain.append(tip)
        link = tip
        while link != root:
            if link not in self.parent_map:
                raise KeyError(f"link '{tip}' is not a descendant of '{root}'")
            jname, parent = self.parent_map[link]
            if joints:
                chain.append(jname)
            if links:
                chain.append(parent)
            link = parent
        chain.reverse()
        return chain
