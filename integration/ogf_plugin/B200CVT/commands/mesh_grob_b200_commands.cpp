/*
 * MeshGrobB200Commands::remesh_smooth — same pre/post-processing and error behaviour as
 * MeshGrobSurfaceCommands::remesh_smooth (src/lib/OGF/mesh/commands/mesh_grob_surface_commands.cpp:363-463);
 * the optimisation itself goes through GEO::remesh_smooth_b200 (integration/geogram_b200.h).
 */
#include <OGF/B200CVT/commands/mesh_grob_b200_commands.h>
#include <geogram_b200.h>
#include <geogram/mesh/mesh_geometry.h>
#include <geogram/mesh/mesh_smoothing.h>

namespace OGF {

    MeshGrob* MeshGrobB200Commands::remesh_smooth(
        const NewMeshGrobName& remesh_name_in, unsigned int nb_points, double tri_shape_adapt, double tri_size_adapt, bool adjust,
        double adjust_max_edge_distance, double adjust_border_importance, unsigned int normal_iter, unsigned int Lloyd_iter,
        unsigned int Newton_iter, unsigned int Newton_m, unsigned int LFS_samples
    ) {
        const std::string remesh_name = remesh_name_in;
        MeshGrob* M = mesh_grob();
        /* user-level problems: Logger::err + nullptr, like the reference command (:379-397) */
        if(remesh_name == M->name()) {
            Logger::err("Remesh") << "remesh should not be the same as mesh" << std::endl;
            return nullptr;
        }
        if(M->facets.nb() == 0) {
            Logger::err("Remesh") << "mesh has no facet" << std::endl;
            return nullptr;
        }
        if(!M->facets.are_simplices()) {
            Logger::err("Remesh") << "mesh need to be simplicial, use repair" << std::endl;
            return nullptr;
        }
        const index_t dimension_before = M->vertices.dimension();
        MeshGrob* remesh = MeshGrob::find_or_create(scene_graph(), remesh_name);
        remesh->clear();
        remesh->lock_graphics();

        if(tri_shape_adapt != 0.0) {
            /* 6D lifting: normals scaled by 0.02 * tri_shape_adapt * bbox diagonal (mesh_geometry.cpp:237-250) */
            GEO::compute_normals(*M);
            M->update();
            if(normal_iter != 0) {
                GEO::simple_Laplacian_smooth(*M, normal_iter, true);
            }
            GEO::set_anisotropy(*M, tri_shape_adapt * 0.02);
        } else {
            M->vertices.set_dimension(3);
        }
        M->update();

        if(tri_size_adapt != 0.0) {
            GEO::compute_sizing_field_b200(*M, tri_size_adapt, LFS_samples);
        } else if(M->vertices.attributes().is_defined("weight")) {
            M->vertices.attributes().delete_attribute_store("weight");
        }
        M->update();

        GEO::remesh_smooth_b200(
            *M, *remesh, nb_points, 0, Lloyd_iter, Newton_iter, Newton_m, adjust, adjust_max_edge_distance, adjust_border_importance
        );

        show_mesh(remesh);
        remesh->unlock_graphics();
        remesh->update();
        if(M->vertices.dimension() != dimension_before) {
            M->vertices.set_dimension(dimension_before);
        }
        M->update();
        return remesh;
    }
}
