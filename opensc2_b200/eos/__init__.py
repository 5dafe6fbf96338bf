"""Coolant table generator: offline host tool that writes the `[10][n_p][n_t]` property tables
the device interpolates in (`osc2_set_fluid_table`), from multiparameter Helmholtz equations of state and the
transport correlations CoolProp 6.4.1 uses for the same fluids.  Nothing here runs per time step."""
from .helmholtz import HelmholtzFluid  # noqa: F401
from .nitrogen import NITROGEN  # noqa: F401
from .tables import build_table, interpolation_error  # noqa: F401
