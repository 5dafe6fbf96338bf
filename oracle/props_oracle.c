/* placeholder: property-fit oracle lands with the property kernels */
