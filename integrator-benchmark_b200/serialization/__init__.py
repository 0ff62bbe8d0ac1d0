from .system_xml import load_system_xml, system_to_xml  # noqa: F401
from .mdtraj_hdf5 import load_mdtraj_hdf5, save_mdtraj_hdf5, topology_from_description  # noqa: F401
