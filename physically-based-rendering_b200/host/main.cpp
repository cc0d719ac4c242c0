/*
 * pbr_headless -- the reference's main() (main.cpp:8-29) without Qt: load config.json, import a
 * model, render frames through PathTracer::generateImage, write the accumulated image.
 *
 *   pbr_headless [--config config.json] --model <dir/> <file.obj> [--frames N] [--out image.pfm]
 *                [--checkpoint acc.bin] [--resume acc.bin] [--deterministic] [--device D]
 *                [--set key=value]...
 */
#include <chrono>
#include <locale.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/pbr_host.h"
#include "Cfg.h"
#include "Logger.h"
#include "qt/GLWidget.h"

int main( int argc, char** argv ) {
	setlocale( LC_ALL, "C" );

	std::string config, dir, file, out, checkpoint, resume;
	std::vector<std::string> sets;
	int frames = 1, device = -1;
	bool deterministic = false;

	for( int i = 1; i < argc; i++ ) {
		const std::string a = argv[i];
		if( a == "--config" && i + 1 < argc ) { config = argv[++i]; }
		else if( a == "--model" && i + 2 < argc ) { dir = argv[++i]; file = argv[++i]; }
		else if( a == "--frames" && i + 1 < argc ) { frames = atoi( argv[++i] ); }
		else if( a == "--out" && i + 1 < argc ) { out = argv[++i]; }
		else if( a == "--checkpoint" && i + 1 < argc ) { checkpoint = argv[++i]; }
		else if( a == "--resume" && i + 1 < argc ) { resume = argv[++i]; }
		else if( a == "--device" && i + 1 < argc ) { device = atoi( argv[++i] ); }
		else if( a == "--set" && i + 1 < argc ) { sets.push_back( argv[++i] ); }
		else if( a == "--deterministic" ) { deterministic = true; }
		else {
			fprintf( stderr, "usage: %s [--config config.json] --model <dir/> <file.obj> [--frames N] [--out image.pfm]\n"
				"       [--checkpoint acc.bin] [--resume acc.bin] [--deterministic] [--device D] [--set key=value]...\n", argv[0] );
			return 2;
		}
	}
	if( file.empty() ) {
		fprintf( stderr, "no model given (--model <dir/> <file.obj>)\n" );
		return 2;
	}

	try {
		if( !config.empty() ) { Cfg::get().loadConfigFile( config.c_str() ); }
		for( size_t i = 0; i < sets.size(); i++ ) {
			const size_t eq = sets[i].find( '=' );
			if( eq == std::string::npos ) { continue; }
			Cfg::get().value<std::string>( sets[i].substr( 0, eq ).c_str(), sets[i].substr( eq + 1 ) );
		}
	}
	catch( const std::exception& e ) {
		fprintf( stderr, "%s\n", e.what() );
		return 1;
	}

	CL::setDefaultDevice( device );
	GLWidget widget;
	widget.getPathTracer()->setDeterministicSeeds( deterministic );
	widget.loadModel( dir, file );

	PathTracer* pt = widget.getPathTracer();
	const int W = (int) pt->getWidth(), H = (int) pt->getHeight();
	std::vector<float> image( (size_t) W * H * 4, 0.0f );

	if( !resume.empty() ) {
		unsigned sc = 0;
		if( pbrh_read_checkpoint( resume.c_str(), image.data(), W, H, &sc ) != 0 ) {
			fprintf( stderr, "%s\n", pbrh_last_error() );
			return 1;
		}
		pt->writeImage( image.data(), sc );
	}

	const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	if( frames > 1 ) { pt->renderFrames( (cl_uint) ( frames - 1 ) ); }
	if( frames > 0 ) { pt->generateImageInto( image.data(), NULL ); }
	const double sec = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();

	uint64_t stats[6];
	pt->getCL()->getStats( stats, false );
	fprintf( stderr, "%d frame(s) %dx%d in %.3f s: %.2f Msamples/s, %.2f Mrays/s, %u samples accumulated\n",
		frames, W, H, sec, frames * (double) W * H * Cfg::get().value<int>( Cfg::RENDER_SAMPLES ) / sec * 1e-6,
		(double) ( stats[0] + stats[1] ) / sec * 1e-6, pt->getSampleCount() );

	if( !out.empty() && pbrh_write_pfm( out.c_str(), image.data(), W, H ) != 0 ) {
		fprintf( stderr, "%s\n", pbrh_last_error() );
		return 1;
	}
	if( !checkpoint.empty() && pbrh_write_checkpoint( checkpoint.c_str(), image.data(), W, H, pt->getSampleCount() ) != 0 ) {
		fprintf( stderr, "%s\n", pbrh_last_error() );
		return 1;
	}
	return 0;
}
