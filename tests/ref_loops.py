"""Caller-side loops of the reference, restated line for line so that the tests can run them against a stand-in of the
third-party object they drive.  TEST INFRASTRUCTURE."""
import numpy as np
import pandas as pd


def compute_rdf_with(RadialDistributionFunction, chemical_symbols, trajectory, dr, rmax):
    """The body of amof.rdf.Rdf.compute_rdf (/root/reference/amof/rdf.py:67-114) with the asap3 class passed in."""
    atomic_numbers_unique = list(set(trajectory[0].get_atomic_numbers()))
    N_species = len(atomic_numbers_unique)

    rmax_half_cell = np.min([a for t in trajectory for a in t.get_cell_lengths_and_angles()[0:3]]) / 2
    if rmax == 'half_cell':
        rmax = rmax_half_cell
    elif rmax > rmax_half_cell:
        rmax = rmax_half_cell

    bins = int(rmax // dr)
    r = np.arange(bins) * dr
    data = pd.DataFrame({"r": r})

    RDFobj = None
    for atoms in trajectory:
        if RDFobj is None:
            RDFobj = RadialDistributionFunction(atoms, rmax, bins)
        else:
            RDFobj.atoms = atoms
        RDFobj.update()

    rdf = RDFobj.get_rdf(groups=0)
    data["X-X"] = rdf

    elements = [[(x, y) for y in atomic_numbers_unique] for x in atomic_numbers_unique]
    partial_rdf = [[0 for y in atomic_numbers_unique] for x in atomic_numbers_unique]
    for i in range(N_species):
        for j in range(N_species):
            xx = elements[i][j]
            xx_str = chemical_symbols[xx[0]] + "-" + chemical_symbols[xx[1]]
            partial_rdf[i][j] = RDFobj.get_rdf(elements=xx, groups=0)
            data[xx_str] = partial_rdf[i][j]
    for i in range(N_species):
        xx = elements[i][i]
        data[chemical_symbols[xx[0]] + "-X"] = sum([partial_rdf[i][j] for j in range(N_species)])
    return data
