/*
 * Graphite plugin "B200CVT": the OGF command of the accelerated path.
 * Same slot signature as MeshGrobSurfaceCommands::remesh_smooth
 * (src/lib/OGF/mesh/commands/mesh_grob_surface_commands.h:180-193), so that Lua / Python scripts only change
 * the interface name:   mesh.I.B200.remesh_smooth(...)   instead of   mesh.I.Surface.remesh_smooth(...)
 *
 * NOT compiled in this repository: gom_class / gom_slots need Graphite's own gomgen step
 * (plugins/OGF/WarpDrive/CMakeLists.txt:60-75 is the pattern). Drop this directory under plugins/OGF/ of a Graphite tree.
 */
#ifndef H_OGF_B200CVT_COMMANDS_MESH_GROB_B200_COMMANDS_H
#define H_OGF_B200CVT_COMMANDS_MESH_GROB_B200_COMMANDS_H

#include <OGF/B200CVT/common/common.h>
#include <OGF/mesh/commands/mesh_grob_commands.h>

namespace OGF {

    gom_class B200CVT_API MeshGrobB200Commands : public MeshGrobCommands {
    public:
        MeshGrobB200Commands() { }
        ~MeshGrobB200Commands() override { }

    gom_slots:
        /**
         * \brief Remeshes a (smooth) surface on the GPU (B200 CVT path).
         * \param[in] remesh name of the generated surface
         * \param[in] nb_points desired number of points in the generated mesh
         * \param[in] tri_shape_adapt adapt triangle shapes (0.0 means no adapt, 1.0 for moderate adapt, ...)
         * \param[in] tri_size_adapt adapt triangle sizes (0.0 means no adapt, 1.0 for moderate adapt, ...)
         * \param[in] adjust if set, vertices are moved to minimize distance with the input surface
         * \advanced
         * \param[in] adjust_max_edge_distance maximum adjustment, relative to local edge length
         * \param[in] adjust_border_importance importance of the fitting term on the borders
         * \param[in] normal_iter number of normal smoothing iterations (if anisotropy is non-zero)
         * \param[in] Lloyd_iter number of Lloyd iterations
         * \param[in] Newton_iter number of Newton iterations
         * \param[in] Newton_m number of Newton evaluations per step
         * \param[in] LFS_samples number of samples (used if gradation is non-zero)
         * \menu /Surface/Remesh
         */
        MeshGrob* remesh_smooth(
            const NewMeshGrobName& remesh = "remesh",
            unsigned int nb_points = 30000,
            double tri_shape_adapt = 1.0,
            double tri_size_adapt = 0.0,
            bool adjust = true,
            double adjust_max_edge_distance = 0.5,
            double adjust_border_importance = 2.0,
            unsigned int normal_iter = 3,
            unsigned int Lloyd_iter = 5,
            unsigned int Newton_iter = 30,
            unsigned int Newton_m = 7,
            unsigned int LFS_samples = 10000
        );
    };
}
#endif
