from .mts_utilities import (generate_all_BAOAB_permutation_strings, generate_baoab_mts_string,  # noqa: F401
                            generate_baoab_mts_string_from_ratios, generate_gbaoab_solvent_solute_string,
                            generate_gbaoab_string, generate_sequential_BAOAB_string)
from .openmm_utilities import (get_masses, get_sum_of_masses, repartition_hydrogen_mass,  # noqa: F401
                               repartition_hydrogen_mass_amber)
