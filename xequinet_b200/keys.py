"""String contract of the data / result dictionaries (same values as the reference's
xequinet/keys.py:4-40, so dictionaries are interchangeable)."""
POSITIONS = "pos"
ATOMIC_NUMBERS = "atomic_numbers"
EDGE_INDEX = "edge_index"
CELL_OFFSETS = "cell_offsets"
CELL = "cell"
PBC = "pbc"
BATCH = "batch"
BATCH_PTR = "ptr"
NUM_GRAPHS = "num_graphs"

CENTER_IDX = 0
NEIGHBOR_IDX = 1

NODE_INVARIANT = "node_invariant"
NODE_EQUIVARIANT = "node_equivariant"  # held in the component-major layout (include/xeq_b200.h)

ATOMIC_ENERGIES = "atomic_energies"
TOTAL_ENERGY = "energy"
FORCES = "forces"
VIRIAL = "virial"

# private entries this package adds to the data dict
GRAPH = "_xeq_graph"
RBF_FREQ = "_xeq_rbf_freq"
RBF_CUTOFF = "_xeq_rbf_cutoff"
STRAIN = "strain"  # nn/basic.py:133-139
POS_EFF = "_xeq_pos_strained"    # positions / cell with the (zero) strain of the virial computation applied:
CELL_EFF = "_xeq_cell_strained"  # what the edge kernels differentiate (nn/basic.py:99-107)
HALO = "_xeq_halo"  # domain.HaloPlan of a spatially sharded run (xequinet_b200/domain.py)

# optional inputs / heads (xequinet/keys.py:44-57, 78-82)
TOTAL_CHARGE = "charge"
TOTAL_SPIN = "spin"
ATOMIC_CHARGES = "atomic_charges"
DIPOLE = "dipole"
DIPOLE_MAGNITUDE = "dipole_magnitude"
POLARIZABILITY = "polarizability"
ISO_POLARIZABILITY = "iso_polarizability"
SPATIAL_EXTENT = "spatial_extent"
SCALAR_OUTPUT = "scalar_output"
CARTESIAN_TENSOR = "cartesian_tensor"

# unit systems of the MD engines the deployment wrappers talk to (xequinet/keys.py:101-120)
LAMMPS_UNIT_STYLE = {
    "metal": {TOTAL_ENERGY: "eV", POSITIONS: "Angstrom", FORCES: "eV/Angstrom", TOTAL_CHARGE: "e"},
    "real": {TOTAL_ENERGY: "kcal/mol", POSITIONS: "Angstrom", FORCES: "kcal/mol/Angstrom", TOTAL_CHARGE: "e"},
    "electron": {TOTAL_ENERGY: "Hartree", POSITIONS: "Bohr", FORCES: "Hartree/Bohr", TOTAL_CHARGE: "e"},
}
