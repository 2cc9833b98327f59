// C ABI glue: handle lifecycle, grid/system/physics setters, state vectors, assembly entry points, instrumentation.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "vfvm_internal.h"

#define NEED(h, cond, msg) \
    if (!(h) || !(cond)) return vfvm_fail((h), (h) ? VFVM_ERR_STATE : VFVM_ERR_ARG, (msg));

extern "C" int vfvm_abi_version(void) { return 1; }

extern "C" int vfvm_create(int device, vfvm_handle** out) {
    if (!out) return VFVM_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return VFVM_ERR_CUDA;  // no CPU fallback: fail loudly
    vfvm_handle* h = new vfvm_handle();
    h->device = device;
    try {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&h->ev0));
        CK(cudaEventCreate(&h->ev1));
        CK(cudaEventCreate(&h->ev2));
        CK(cudaEventCreate(&h->ev3));
        CK(cudaEventCreate(&h->ev4));
        CK(cudaMallocHost((void**)&h->flags_host, 4 * sizeof(int32_t)));
        CK(cudaMallocHost((void**)&h->red_host, 8 * sizeof(double)));
        h->flags.alloc(4);
        h->red.alloc(32);
        CK(cudaMemset(h->flags.p, 0, 4 * sizeof(int32_t)));
        CK(cudaMemset(h->red.p, 0, 32 * sizeof(double)));
        memset(&h->phys, 0, sizeof(h->phys));
        memset(&h->masks, 0, sizeof(h->masks));
    } catch (const std::string& msg) {
        delete h;
        return VFVM_ERR_CUDA;
    }
    // byte accounting for the long-lived buffers
    DevBuf<double>* dbl[] = {&h->coord, &h->nf_fac, &h->ef_fac, &h->bfacenodefac, &h->nzfac, &h->offval, &h->diagval, &h->vec[0], &h->vec[1], &h->vec[2], &h->vec[3],
                             &h->pc_diag, &h->ilu_off, &h->ilu_diag, &h->nodal_source, &h->send_buf, &h->node_q};
    for (auto* b : dbl) b->tally = &h->bytes;
    for (auto& w : h->work) w.tally = &h->bytes;
    DevBuf<int32_t>* i32[] = {&h->cellnodes, &h->cellregions, &h->bfacenodes, &h->bfaceregions, &h->edgenodes, &h->celledges, &h->nf_region, &h->ef_region,
                              &h->rowptr, &h->colidx, &h->sell_ptr, &h->nz_edge, &h->tile_row, &h->bn_node, &h->bn_ptr, &h->bn_bface, &h->bn_local, &h->upos, &h->ilu_rank, &h->ilu_lrows, &h->ilu_urows, &h->send_idx};
    for (auto* b : i32) b->tally = &h->bytes;
    h->nf_colptr.tally = h->ef_colptr.tally = &h->bytes;
    h->offval32.tally = &h->bytes;
    *out = h;
    return VFVM_OK;
}

int vfvm_comm_destroy(vfvm_handle* h);
int vfvm_halo_exchange_ptr(vfvm_handle* h, double* x);

extern "C" void vfvm_destroy(vfvm_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    vfvm_comm_destroy(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    vfvm_amg_free(h);
    if (h->iter_graph) cudaGraphExecDestroy((cudaGraphExec_t)h->iter_graph);
    if (h->flags_host) cudaFreeHost(h->flags_host);
    if (h->red_host) cudaFreeHost(h->red_host);
    for (cudaEvent_t e : h->pipe_ev) cudaEventDestroy(e);
    if (h->stream_in) cudaStreamDestroy(h->stream_in);
    if (h->stream_out) cudaStreamDestroy(h->stream_out);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev2);
    cudaEventDestroy(h->ev3);
    cudaEventDestroy(h->ev4);
    cudaStream_t s = h->stream;
    delete h;
    if (s) cudaStreamDestroy(s);
}

extern "C" const char* vfvm_last_error(vfvm_handle* h) { return h ? h->err.c_str() : "null handle (vfvm_create failed: no CUDA device?)"; }

extern "C" int vfvm_set_grid(vfvm_handle* h, int dim, int coordsys, int64_t nnodes, int64_t ncells, int64_t nbfaces, const double* coord,
                             const int32_t* cellnodes, const int32_t* cellregions, const int32_t* bfacenodes, const int32_t* bfaceregions) {
    if (!h) return VFVM_ERR_ARG;
    if (dim < 1 || dim > 3 || nnodes <= 0 || ncells <= 0 || nbfaces < 0 || !coord || !cellnodes || !cellregions) return vfvm_fail(h, VFVM_ERR_ARG, "bad grid arguments");
    if (coordsys < VFVM_CARTESIAN || coordsys > VFVM_SPHERICAL || (coordsys == VFVM_SPHERICAL && dim != 1) || (coordsys == VFVM_CYLINDRICAL && dim == 3))
        return vfvm_fail(h, VFVM_ERR_ARG, "coordinate system not available in this space dimension (src/vfvm_xgrid.jl:23-47)");
    if (nnodes >= ((int64_t)1 << 31)) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "more than 2^31 nodes");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        h->dim = dim;
        h->coordsys = coordsys;
        h->N = nnodes;
        h->Nown = nnodes;
        h->C = ncells;
        h->NB = nbfaces;
        int rmin = INT32_MAX, rmax = 0, brmax = 0;
        for (int64_t c = 0; c < ncells; c++) {
            rmin = std::min(rmin, cellregions[c]);
            rmax = std::max(rmax, cellregions[c]);
        }
        for (int64_t b = 0; b < nbfaces; b++) brmax = std::max(brmax, bfaceregions[b]);
        if (rmin < 1) return vfvm_fail(h, VFVM_ERR_ARG, "cell region labels must be >= 1");
        for (int64_t i = 0; i < ncells * (dim + 1); i++)
            if (cellnodes[i] < 0 || cellnodes[i] >= nnodes) return vfvm_fail(h, VFVM_ERR_ARG, "cellnodes entry out of range (indices are 0-based)");
        for (int64_t i = 0; i < nbfaces * dim; i++)
            if (bfacenodes[i] < 0 || bfacenodes[i] >= nnodes) return vfvm_fail(h, VFVM_ERR_ARG, "bfacenodes entry out of range (indices are 0-based)");
        h->ncellregions = rmax;
        h->nbfaceregions = brmax;
        h->single_region = (rmin == rmax);
        h->the_region = rmin;
        h->coord.upload(coord, (size_t)dim * nnodes, h->stream);
        h->cellnodes.upload(cellnodes, (size_t)(dim + 1) * ncells, h->stream);
        h->cellregions.upload(cellregions, ncells, h->stream);
        h->bfacenodes.upload(bfacenodes, (size_t)dim * nbfaces, h->stream);
        h->bfaceregions.upload(bfaceregions, nbfaces, h->stream);
        CK(cudaStreamSynchronize(h->stream));
        h->have_grid = true;
        h->have_geometry = h->have_pattern = false;
    })
    return VFVM_OK;
}

extern "C" int vfvm_set_owned_nodes(vfvm_handle* h, int64_t n_owned) {
    NEED(h, h->have_grid, "vfvm_set_grid has not been called");
    if (n_owned <= 0 || n_owned > h->N) return vfvm_fail(h, VFVM_ERR_ARG, "n_owned out of range");
    h->Nown = n_owned;
    h->have_pattern = false;
    return VFVM_OK;
}

extern "C" int vfvm_build_geometry(vfvm_handle* h) {
    NEED(h, h->have_grid, "vfvm_set_grid has not been called");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        return vfvm_geometry_build(h);
    })
}

extern "C" int vfvm_num_edges(vfvm_handle* h, int64_t* nedges) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    *nedges = h->E;
    return VFVM_OK;
}
extern "C" int vfvm_get_edgenodes(vfvm_handle* h, int32_t* out) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    h->edgenodes.download(out, h->stream);
    return VFVM_OK;
}
extern "C" int vfvm_get_celledges(vfvm_handle* h, int32_t* out) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    h->celledges.download(out, h->stream);
    return VFVM_OK;
}
extern "C" int vfvm_num_factors(vfvm_handle* h, int64_t* nnf, int64_t* nef) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    *nnf = (int64_t)h->nf_fac.n;
    *nef = (int64_t)h->ef_fac.n;
    return VFVM_OK;
}
extern "C" int vfvm_get_nodefactors(vfvm_handle* h, int64_t* colptr, int32_t* region, double* fac) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    h->nf_colptr.download(colptr, h->stream);
    h->nf_region.download(region, h->stream);
    h->nf_fac.download(fac, h->stream);
    return VFVM_OK;
}
extern "C" int vfvm_get_edgefactors(vfvm_handle* h, int64_t* colptr, int32_t* region, double* fac) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    h->ef_colptr.download(colptr, h->stream);
    h->ef_region.download(region, h->stream);
    h->ef_fac.download(fac, h->stream);
    return VFVM_OK;
}
extern "C" int vfvm_get_bfacefactors(vfvm_handle* h, double* out) {
    NEED(h, h->have_geometry, "vfvm_build_geometry has not been called");
    h->bfacenodefac.download(out, h->stream);
    return VFVM_OK;
}

extern "C" int vfvm_set_system(vfvm_handle* h, int nspecies, const uint8_t* region_species) {
    NEED(h, h->have_grid, "vfvm_set_grid has not been called");
    if (nspecies < 1 || nspecies > VFVM_MAX_SPECIES) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "number of species out of range (1..10)");
    if (!(nspecies <= 5 || nspecies == 10)) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "species counts with a device instantiation: 1,2,3,4,5,10");
    if (h->ncellregions > VFVM_MAX_CREGIONS) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "more than 32 cell regions");
    // region_species (n x ncellregions, column-major): species enabled per cell region, enable_species! src/vfvm_system.jl:433-480
    h->region_species.assign((size_t)nspecies * h->ncellregions, 1);
    h->masked = false;
    if (region_species)
        for (int i = 0; i < nspecies * h->ncellregions; i++) {
            h->region_species[i] = region_species[i] ? 1 : 0;
            if (!region_species[i]) h->masked = true;
        }
    if (h->nbfaceregions > VFVM_MAX_BREGIONS) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "more than 16 boundary regions");
    h->n = nspecies;
    h->bregion_species.clear();
    h->nbregions_bs = 0;
    PhysicsDev& ph = h->phys;
    memset(&ph, 0, sizeof(ph));
    ph.nbregions = h->nbfaceregions;
    h->phys_dirty = true;
    h->have_system = true;
    h->have_pattern = false;
    return VFVM_OK;
}

extern "C" int vfvm_set_boundary_species(vfvm_handle* h, int nbregions, const uint8_t* bregion_species) {
    NEED(h, h->have_system, "vfvm_set_system has not been called");
    if (nbregions < 0 || nbregions > VFVM_MAX_BREGIONS) return vfvm_fail(h, VFVM_ERR_ARG, "more than 16 boundary regions");
    h->bregion_species.clear();
    h->nbregions_bs = 0;
    bool any = false;
    if (bregion_species)
        for (int i = 0; i < h->n * nbregions; i++) any |= bregion_species[i] != 0;
    if (any) {
        for (int r = 0; r < nbregions; r++)
            for (int i = 0; i < h->n; i++)
                if (bregion_species[r * h->n + i])
                    for (int c = 0; c < h->ncellregions; c++)
                        if (h->region_species[(size_t)c * h->n + i]) return vfvm_fail(h, VFVM_ERR_ARG, "species is already a bulk species (src/vfvm_system.jl:496-498)");
        h->bregion_species.assign(bregion_species, bregion_species + (size_t)h->n * nbregions);
        h->nbregions_bs = nbregions;
        h->masked = true;  // the boundary species is undefined in the interior: identity rows there
    }
    h->have_pattern = false;
    return VFVM_OK;
}

static int min_species(int slot, int id) {
    if (slot == VFVM_SLOT_FLUX) {
        if (id == VFVM_FLUX_CROSSDIFF2 || id == VFVM_FLUX_SG_UNIPOLAR || id == VFVM_FLUX_SEDAN) return 2;
        if (id == VFVM_FLUX_SG_BIPOLAR) return 3;
        if (id == VFVM_FLUX_MIXTURE) return 2;
    }
    if (slot == VFVM_SLOT_REACTION) {
        if (id == VFVM_REACTION_BILINEAR2) return 2;
        if (id == VFVM_REACTION_BIPOLAR) return 3;
    }
    if (slot == VFVM_SLOT_STORAGE && id == VFVM_STORAGE_BIPOLAR) return 3;
    if (slot == VFVM_SLOT_BREACTION && id == VFVM_BREACTION_CATALYSIS) return 3;
    if (slot == VFVM_SLOT_EDGEREACTION && id == VFVM_EDGEREACTION_JOULE) return 2;
    return 1;
}

extern "C" int vfvm_set_physics(vfvm_handle* h, int slot, int id, const double* params, int np) {
    NEED(h, h->have_system, "vfvm_set_system has not been called");
    if (slot < 0 || slot >= VFVM_NUM_SLOTS || np < 0 || (np > 0 && !params)) return vfvm_fail(h, VFVM_ERR_ARG, "bad slot / params");
    static const int maxid[VFVM_NUM_SLOTS] = {VFVM_FLUX_MIXTURE, VFVM_REACTION_REGION_AFFINE, VFVM_STORAGE_BIPOLAR, VFVM_SOURCE_NODAL, VFVM_BREACTION_POW,
                                              VFVM_EDGEREACTION_JOULE, VFVM_BSTORAGE_LINEAR};
    if (id < 0 || id > maxid[slot])
        return vfvm_fail(h, VFVM_ERR_UNREGISTERED, "physics id is not in the registered device library; arbitrary host callbacks are not evaluated (no CPU fallback)");
    const int n = h->n;
    if (n < min_species(slot, id)) return vfvm_fail(h, VFVM_ERR_ARG, "this physics id needs more species than the system has");
    // exact-size checks and device restrictions
    int need = -1;
    if (slot == VFVM_SLOT_FLUX) {
        const int t[] = {0, n, n + 1, 3, 3, 5, 10, n + n * n};
        need = t[id];
        if (id == VFVM_FLUX_MIXTURE && !(n == 2 || n == 3 || n == 5)) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "mixture flux: device instantiations for 2, 3 and 5 species");
        if ((id == VFVM_FLUX_CROSSDIFF2 || id == VFVM_FLUX_SG_UNIPOLAR || id == VFVM_FLUX_SEDAN) && n != 2)
            return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "this flux has a device instantiation for exactly 2 species");
        if (id == VFVM_FLUX_SG_BIPOLAR && (n != 3 || np != 10 || params[7] != 0 || params[8] != 1 || params[9] != 2))
            return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "bipolar SG flux: device instantiation needs 3 species ordered (iphin, iphip, ipsi) = (1,2,3)");
    } else if (slot == VFVM_SLOT_REACTION) {
        if (id == VFVM_REACTION_POW) need = 2 * n;
        if (id == VFVM_REACTION_SINH) need = n;
        if (id == VFVM_REACTION_AFFINE) need = n * n + n;
        if (id == VFVM_REACTION_BILINEAR2) need = 1;
        if (id == VFVM_REACTION_REGION_AFFINE) {
            if (np < 1 || (int)params[0] < 1 || np != 1 + (int)params[0] * (n * n + n)) return vfvm_fail(h, VFVM_ERR_ARG, "region-affine reaction: params = nreg, then nreg x (R[n*n], r0[n])");
            need = np;
        }
        if (id == VFVM_REACTION_BILINEAR2 && n != 2) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "bilinear reaction: exactly 2 species");
        if (id == VFVM_REACTION_BIPOLAR) {
            if (n != 3 || np < 9 || params[5] != 0 || params[6] != 1 || params[7] != 2 || np != 9 + (int)params[8])
                return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "bipolar reaction: 3 species ordered (iphin, iphip, ipsi) = (1,2,3)");
            need = np;
        }
    } else if (slot == VFVM_SLOT_STORAGE) {
        if (id == VFVM_STORAGE_LINEAR) need = n;
        if (id == VFVM_STORAGE_POW) need = 2 * n;
        if (id == VFVM_STORAGE_BIPOLAR) {
            need = 7;
            if (n != 3 || np != 7 || params[4] != 0 || params[5] != 1 || params[6] != 2)
                return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "bipolar storage: 3 species ordered (iphin, iphip, ipsi) = (1,2,3)");
        }
    } else if (slot == VFVM_SLOT_SOURCE) {
        const int t[] = {0, n, 5, 2, 4, 2 * n, 0};
        need = t[id];
    } else if (slot == VFVM_SLOT_BREACTION) {
        if (id == VFVM_BREACTION_LINEAR) need = 1 + n * n;
        if (id == VFVM_BREACTION_POW) need = 1 + 2 * n;
        if (id == VFVM_BREACTION_CATALYSIS) {
            need = 9;
            if (np == 9)
                for (int k = 6; k < 9; k++)
                    if (params[k] < 0 || params[k] >= n) return vfvm_fail(h, VFVM_ERR_ARG, "catalysis boundary reaction: species index out of range");
        }
    } else if (slot == VFVM_SLOT_EDGEREACTION) {
        if (id == VFVM_EDGEREACTION_DIAMOND) need = n;
        if (id == VFVM_EDGEREACTION_JOULE) {
            need = 3;
            if (np == 3 && (params[1] < 0 || params[1] >= n || params[2] < 0 || params[2] >= n)) return vfvm_fail(h, VFVM_ERR_ARG, "Joule heat edge reaction: species index out of range");
        }
    } else if (slot == VFVM_SLOT_BSTORAGE) {
        if (id == VFVM_BSTORAGE_LINEAR) need = 1 + n;
    }
    if (id == VFVM_NONE) need = np;  // params ignored
    if (need >= 0 && np != need) return vfvm_fail(h, VFVM_ERR_ARG, "parameter block has the wrong length for this physics id");
    // (re)pack the parameter blocks
    PhysicsDev& ph = h->phys;
    std::vector<double> blocks[VFVM_NUM_SLOTS];
    for (int s = 0; s < VFVM_NUM_SLOTS; s++) blocks[s].assign(ph.params + ph.slot[s].off, ph.params + ph.slot[s].off + ph.slot[s].np);
    blocks[slot].assign(params, params + (id == VFVM_NONE ? 0 : np));
    int off = 0;
    for (int s = 0; s < VFVM_NUM_SLOTS; s++) off += (int)blocks[s].size();
    if (off > VFVM_MAX_PARAMS) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "parameter blocks exceed 160 doubles");
    const bool structure_change = (ph.slot[slot].id != id);
    ph.slot[slot].id = id;
    off = 0;
    for (int s = 0; s < VFVM_NUM_SLOTS; s++) {
        ph.slot[s].off = off;
        ph.slot[s].np = (int)blocks[s].size();
        std::copy(blocks[s].begin(), blocks[s].end(), ph.params + off);
        off += ph.slot[s].np;
    }
    h->phys_dirty = true;
    if (h->have_pattern) {
        // a parameter change that alters a coupling mask (a coefficient becoming zero / nonzero) needs a new pattern
        Masks old = h->masks;
        vfvm_physics_masks(h);
        if (structure_change || memcmp(&old, &h->masks, sizeof(Masks)) != 0) h->have_pattern = false;
    }
    return VFVM_OK;
}

extern "C" int vfvm_set_nodal_source(vfvm_handle* h, const double* table) {
    NEED(h, h->have_system, "vfvm_set_system has not been called");
    VFVM_TRY(h, {
        h->nodal_source.upload(table, (size_t)h->n * h->N, h->stream);
        CK(cudaStreamSynchronize(h->stream));
        h->phys.nodal_source = h->nodal_source.p;
        h->phys_dirty = true;
    })
    return VFVM_OK;
}

extern "C" int vfvm_set_legacy_bc(vfvm_handle* h, int nbregions, const double* factors, const double* values) {
    NEED(h, h->have_system, "vfvm_set_system has not been called");
    // a rank-local grid piece may not contain faces of every boundary region: the table may be larger than the local maximum
    if (nbregions < h->nbfaceregions || nbregions > VFVM_MAX_BREGIONS) return vfvm_fail(h, VFVM_ERR_ARG, "nbregions does not match the grid (or exceeds 16)");
    PhysicsDev& ph = h->phys;
    bool any = false;
    for (int i = 0; i < h->n * nbregions; i++) {
        ph.bfactors[i] = factors[i];
        ph.bvalues[i] = values[i];
        any |= (factors[i] != 0.0) || (values[i] != 0.0);
    }
    ph.has_legacy_bc = any ? 1 : 0;  // src/vfvm_assembly.jl:329
    ph.nbregions = nbregions;
    h->phys_dirty = true;
    if (h->have_pattern) {
        Masks old = h->masks;
        vfvm_physics_masks(h);
        if (memcmp(&old, &h->masks, sizeof(Masks)) != 0) h->have_pattern = false;
    }
    return VFVM_OK;
}

extern "C" int vfvm_set_bc_entries(vfvm_handle* h, int nentries, const vfvm_bc_entry* entries) {
    NEED(h, h->have_system, "vfvm_set_system has not been called");
    if (nentries < 0 || nentries > VFVM_MAX_BC) return vfvm_fail(h, VFVM_ERR_UNSUPPORTED, "more than 32 boundary condition entries");
    for (int i = 0; i < nentries; i++) {
        if (entries[i].kind < VFVM_BC_DIRICHLET || entries[i].kind > VFVM_BC_ROBIN) return vfvm_fail(h, VFVM_ERR_UNREGISTERED, "unknown boundary condition kind");
        if (entries[i].species < 0 || entries[i].species >= h->n) return vfvm_fail(h, VFVM_ERR_ARG, "boundary condition species out of range");
        h->phys.bc[i] = entries[i];
    }
    h->phys.nbc = nentries;
    h->phys_dirty = true;
    if (h->have_pattern) {
        Masks old = h->masks;
        vfvm_physics_masks(h);
        if (memcmp(&old, &h->masks, sizeof(Masks)) != 0) h->have_pattern = false;
    }
    return VFVM_OK;
}

void vfvm_sync_physics(vfvm_handle* h) {
    if (!h->phys_dirty && h->phys_dev.p) return;
    h->phys_dev.upload(&h->phys, 1, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    vfvm_source_cache(h);  // the source callback does not depend on u: tabulate it once per physics change
    CK(cudaStreamSynchronize(h->stream));
    h->phys_dirty = false;
}

extern "C" int vfvm_build_pattern(vfvm_handle* h) {
    NEED(h, h->have_geometry && h->have_system, "vfvm_build_geometry and vfvm_set_system must come first");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        int rc = vfvm_pattern_build(h);
        if (rc) return rc;
        vfvm_sync_physics(h);
    })
    return VFVM_OK;
}

// ---- state vectors -----------------------------------------------------------------------------------------
extern "C" int vfvm_set_vector(vfvm_handle* h, int which, const double* src, int memspace) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    if (which < 0 || which > 3 || !src) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id");
    VFVM_TRY(h, {
        CK(cudaMemcpyAsync(h->vec[which].p, src, sizeof(double) * h->n * h->N, memspace == VFVM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    })
    return VFVM_OK;
}
extern "C" int vfvm_get_vector(vfvm_handle* h, int which, double* dst, int memspace) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    if (which < 0 || which > 3 || !dst) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id");
    VFVM_TRY(h, {
        CK(cudaMemcpyAsync(dst, h->vec[which].p, sizeof(double) * h->n * h->N, memspace == VFVM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    })
    return VFVM_OK;
}
extern "C" int vfvm_copy_vector(vfvm_handle* h, int dst, int src) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    if (dst < 0 || dst > 3 || src < 0 || src > 3) return vfvm_fail(h, VFVM_ERR_ARG, "bad vector id");
    VFVM_TRY(h, {
        if (dst != src) CK(cudaMemcpyAsync(h->vec[dst].p, h->vec[src].p, sizeof(double) * h->n * h->N, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    })
    return VFVM_OK;
}

extern "C" int vfvm_init_dirichlet(vfvm_handle* h, double time, double lambda) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        vfvm_sync_physics(h);
        int rc = vfvm_init_dirichlet_impl(h, time, lambda);
        if (rc == VFVM_OK && h->nranks > 1) {  // halo copies of Dirichlet nodes follow their owners
            vfvm_halo_exchange_ptr(h, h->vec[VFVM_VEC_SOLUTION].p);
            CK(cudaStreamSynchronize(h->stream));
            if (int prc = vfvm_peer_check(h)) return prc;
        }
        return rc;
    })
}

extern "C" int vfvm_assemble(vfvm_handle* h, double time, double tstep, double lambda) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called (or physics structure changed since)");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        vfvm_sync_physics(h);
        return vfvm_assemble_impl(h, time, tstep, lambda);
    })
}

extern "C" int vfvm_assemble_async(vfvm_handle* h, double time, double tstep, double lambda) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called (or physics structure changed since)");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        vfvm_sync_physics(h);
        return vfvm_assemble_impl(h, time, tstep, lambda, true);
    })
}

extern "C" int vfvm_sync(vfvm_handle* h) {
    if (!h) return VFVM_ERR_ARG;
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        return vfvm_assemble_finish(h);
    })
}

extern "C" int vfvm_eval_res_jac(vfvm_handle* h, const double* U, const double* UOld, double* F, int memspace, double time, double tstep, double lambda) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called (or physics structure changed since)");
    if (!U || !F) return vfvm_fail(h, VFVM_ERR_ARG, "null vector");
    VFVM_TRY(h, {
        CK(cudaSetDevice(h->device));
        if (memspace == VFVM_HOST && vfvm_pipeline_applies(h)) {
            vfvm_sync_physics(h);
            return vfvm_eval_res_jac_pipelined(h, U, UOld, F, time, tstep, lambda);
        }
        const size_t bytes = sizeof(double) * h->n * h->N;
        const cudaMemcpyKind in = memspace == VFVM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        const cudaMemcpyKind out = memspace == VFVM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        CK(cudaMemcpyAsync(h->vec[VFVM_VEC_SOLUTION].p, U, bytes, in, h->stream));
        if (UOld && UOld != U) CK(cudaMemcpyAsync(h->vec[VFVM_VEC_OLDSOL].p, UOld, bytes, in, h->stream));
        else CK(cudaMemcpyAsync(h->vec[VFVM_VEC_OLDSOL].p, h->vec[VFVM_VEC_SOLUTION].p, bytes, cudaMemcpyDeviceToDevice, h->stream));
        vfvm_sync_physics(h);
        int rc = vfvm_assemble_impl(h, time, tstep, lambda);
        CK(cudaMemcpyAsync(F, h->vec[VFVM_VEC_RESIDUAL].p, bytes, out, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return rc;
    })
}

// ---- instrumentation -----------------------------------------------------------------------------------------
extern "C" int vfvm_timings(vfvm_handle* h, double* ms_out) {
    if (!h || !ms_out) return VFVM_ERR_ARG;
    for (int i = 0; i < VFVM_NUM_TIMES; i++) ms_out[i] = h->times[i];
    return VFVM_OK;
}
extern "C" int vfvm_launch_count(vfvm_handle* h, int64_t* n) {
    if (!h || !n) return VFVM_ERR_ARG;
    *n = h->launches;
    return VFVM_OK;
}
extern "C" int vfvm_stream(vfvm_handle* h, void** s) {
    if (!h || !s) return VFVM_ERR_ARG;
    *s = (void*)h->stream;
    return VFVM_OK;
}
extern "C" int vfvm_device_bytes(vfvm_handle* h, int64_t* bytes) {
    if (!h || !bytes) return VFVM_ERR_ARG;
    *bytes = h->bytes;
    return VFVM_OK;
}
extern "C" int vfvm_plane_counts(vfvm_handle* h, int* off_planes, int* diag_planes) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    *off_planes = h->cF;
    *diag_planes = h->cD;
    return VFVM_OK;
}
extern "C" int vfvm_block_counts(vfvm_handle* h, int64_t* nblocks_off, int64_t* nblocks_stored) {
    NEED(h, h->have_pattern, "vfvm_build_pattern has not been called");
    *nblocks_off = h->nnz_off;      // off-diagonal blocks of the owned rows = 2 x edges (summed over ranks)
    *nblocks_stored = h->nnz_sell;  // including SELL-32 padding
    return VFVM_OK;
}
