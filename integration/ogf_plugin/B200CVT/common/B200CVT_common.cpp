/*
 * Plugin entry point (pattern: plugins/OGF/WarpDrive/common/WarpDrive_common.cpp:58-96): registers the command class on
 * MeshGrob and the "B200NN" Delaunay backend in geogram's factory.
 */
#include <OGF/B200CVT/common/common.h>
#include <OGF/B200CVT/commands/mesh_grob_b200_commands.h>
#include <OGF/basic/modules/module.h>
#include <OGF/gom/types/gom_defs.h>
#include <OGF/scene_graph/types/scene_graph_library.h>
#include <geogram_b200.h>

namespace OGF {

    void B200CVT_libinit::initialize() {
        Logger::out("Init") << "<B200CVT>" << std::endl;
        gom_package_initialize(B200CVT);
        ogf_register_grob_commands<OGF::MeshGrob, OGF::MeshGrobB200Commands>();
        GEO::b200_register();    /* Delaunay::create(dim, "B200NN"), `algo:delaunay=B200NN` */
        Module* module_info = new Module;
        module_info->set_name("B200CVT");
        module_info->set_vendor("OGF");
        module_info->set_version("3.0");
        module_info->set_info("CVT remeshing on NVIDIA B200 through libb200cvt");
        Module::bind_module("B200CVT", module_info);
        Logger::out("Init") << "</B200CVT>" << std::endl;
    }

    void B200CVT_libinit::terminate() {
        Logger::out("Init") << "<~B200CVT>" << std::endl;
        Module::unbind_module("B200CVT");
        Logger::out("Init") << "</~B200CVT>" << std::endl;
    }
}
