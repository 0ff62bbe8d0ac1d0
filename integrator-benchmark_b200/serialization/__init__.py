from .system_xml import load_system_xml, system_to_xml  # noqa: F401
